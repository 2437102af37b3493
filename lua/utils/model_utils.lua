-- Drop-in for utils/model_utils.lua (clone_list MU:4, clone_many_times MU:15, combine_all_parameters MU:38).
-- Construction-time helpers: no arithmetic crosses the C ABI here.  They rely only on the stock nn.Module
-- protocol (parameters / clone / getParameters), which the librau-backed modules implement through child
-- nn.Linear parameter holders, so weight sharing across unrolled copies behaves as with the nngraph cells.
local model_utils = {}

function model_utils.clone_list(tensor_list, zero_too)
  local out = {}
  for k, v in pairs(tensor_list) do
    out[k] = v:clone()
    if zero_too then out[k]:zero() end
  end
  return out
end

-- T copies whose parameter and gradient tensors alias the prototype's (MU:15-36).  Like the reference this goes
-- through parameters() and Tensor:set, NOT through clone(names...)/share: parameters() walks child modules whatever the
-- container class, so the result does not depend on how share() recurses.
function model_utils.clone_many_times(net, T)
  local clones = {}
  local protoW, protoG
  if net.parameters then protoW, protoG = net:parameters() end
  for t = 1, T do
    local c = net:clone()                        -- torch.MemoryFile round trip; nothing is shared yet
    if protoW then
      local w, g = c:parameters()
      assert(#w == #protoW, 'clone_many_times: the clone lists a different number of parameter tensors')
      for i = 1, #w do
        w[i]:set(protoW[i])
        g[i]:set(protoG[i])
      end
    end
    clones[t] = c
    collectgarbage()
  end
  return clones
end

-- one flat parameter vector and one flat gradient vector over several networks; tensors that already share a
-- storage are laid out once
function model_utils.combine_all_parameters(...)
  local box = nn.Container and nn.Container() or nn.Sequential()
  for _, net in ipairs({...}) do box:add(net) end
  return box:getParameters()
end

return model_utils
