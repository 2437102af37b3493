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

-- T copies whose parameter and gradient tensors alias the prototype's
function model_utils.clone_many_times(net, T)
  local clones = {}
  for t = 1, T do
    clones[t] = net.parameters and net:clone('weight', 'bias', 'gradWeight', 'gradBias') or net:clone()
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
