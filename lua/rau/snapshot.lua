-- lua/rau/snapshot.lua -- the `.t7` snapshots of the experiment scripts for the STEP-level integration (INTEGRATION.md).
--
-- The scripts save {it, opt, epoch, params = {embed_param, rnn_param, mult_param}} (F:1223-1232) and Eval.lua copies the
-- three vectors back into freshly flattened parameters (EV:114, EV:345-347).  torch.save / torch.load do the I/O natively;
-- what the step-level integration needs is the element order: the vectors are getParameters() results (F:322-324), i.e.
-- nngraph's module traversal order, while rau_train_step works on librau's flat layout (rau_param_offset, include/rau.h).
-- `embed` and `rnn` coincide; in `mult` three blocks sit elsewhere.  Same tables as rau_vqa_b200/utils/snapshot.py, where
-- the derivation of nngraph's order from F:231-307 / A:4-74 is written down and tested.  No arithmetic here: narrow + copy.
local M = {}

M.LIBRAU_ORDER = {'Wq','bq','Wh','bh','Wi','bi','Wqa','bqa','Wa','ba','ws','Wm','bm','Wp','bp','Wx','bx','Whh','bhh',
                  'Wo','bo','Ws','bso','wd','bs','bd'}
M.NNGRAPH_ORDER = {'Wq','bq','Wh','bh','Wi','bi','Wa','ba','Wqa','bqa','ws','bs','Wm','bm','Wp','bp','Wx','bx','Whh','bhh',
                   'Wo','bo','Ws','bso','wd','bd'}

-- element counts of the mult group's tensors; c = {Hq=, nlayer=, C=, S=, M=, A=, H=, N=}
function M.mult_sizes(c)
  local Q = 2 * c.Hq * c.nlayer
  return {Wq = c.M * Q, bq = c.M, Wh = c.M * c.H, bh = c.M, Wi = c.M * c.C, bi = c.M, Wqa = c.A * c.M, bqa = c.A,
          Wa = c.A * c.M, ba = c.A, ws = c.A, bs = 1, Wm = c.S * c.H, bm = c.S, Wp = c.M * c.S, bp = c.M,
          Wx = 4 * c.H * c.M, bx = 4 * c.H, Whh = 4 * c.H * c.H, bhh = 4 * c.H, Wo = c.M * c.H, bo = c.M,
          Ws = c.N * c.M, bso = c.N, wd = c.M, bd = 1}
end

local function offsets(order, sizes)
  local off, pos = {}, 0
  for _, name in ipairs(order) do off[name] = pos; pos = pos + sizes[name] end
  return off, pos
end

-- dst (flat vector in `dst_order`) <- src (flat vector in `src_order`), tensor by tensor
local function reorder(src, c, src_order, dst_order)
  local sizes = M.mult_sizes(c)
  local so, total = offsets(src_order, sizes)
  local do_, total2 = offsets(dst_order, sizes)
  assert(total == total2 and src:nElement() == total, 'mult vector length does not match the configuration')
  local dst = src.new():resizeAs(src)
  for name, n in pairs(sizes) do
    dst:narrow(1, do_[name] + 1, n):copy(src:narrow(1, so[name] + 1, n))
  end
  return dst
end

-- snapshot / getParameters() vector of the reference -> the vector rau_train_step expects, and back
function M.mult_from_nngraph(v, c) return reorder(v, c, M.NNGRAPH_ORDER, M.LIBRAU_ORDER) end
function M.mult_to_nngraph(v, c) return reorder(v, c, M.LIBRAU_ORDER, M.NNGRAPH_ORDER) end

-- what Eval.lua does at EV:114 / EV:345-347, for librau's layout
function M.load(path, c)
  local snap = torch.load(path)
  return {it = snap.it, epoch = snap.epoch, opt = snap.opt,
          params = {snap.params[1], snap.params[2], M.mult_from_nngraph(snap.params[3], c)}}
end

-- what F:1223-1232 does
function M.save(path, it, epoch, opt, params, c)
  torch.save(path, {it = it, opt = opt, epoch = epoch,
                    params = {[1] = params[1], [2] = params[2], [3] = M.mult_to_nngraph(params[3], c)}})
end

return M
