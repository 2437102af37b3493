-- Drop-in for utils/optim_updates.lua: the same six GLOBAL functions with the same signatures and the same
-- `state` table fields (m, v, t), each one fused kernel launch (rau_optim_step) on flat torch.CudaTensors
-- instead of ~9 tensor ops (OU:7-87).  No CPU path: CPU tensors raise.
local rau = require 'rau.ffi'
local C = rau.C
local OPT = {sgd = 0, sgdm = 1, sgdmom = 2, adagrad = 3, rmsprop = 4, adam = 5}

local function slot(state, key, like)
  if not state[key] then state[key] = like.new(like:size()):zero() end
  return state[key]
end

local function step(kind, x, dx, lr, h0, h1, h2, s0, s1, t)
  rau.check(C.rau_optim_step(rau.ctx(), kind, x:nElement(), rau.fptr(x), rau.fptr(dx), lr, h0 or 0, h1 or 0, h2 or 0,
                             rau.fptr(s0), rau.fptr(s1), t or 1))
end

function sgd(x, dx, lr) step(OPT.sgd, x, dx, lr) end                                                    -- OU:7-9
function sgdm(x, dx, lr, alpha, state) step(OPT.sgdm, x, dx, lr, alpha, 0, 0, slot(state, 'v', x)) end   -- OU:11-19
function sgdmom(x, dx, lr, alpha, state) step(OPT.sgdmom, x, dx, lr, alpha, 0, 0, slot(state, 'm', x)) end   -- OU:21-31
function adagrad(x, dx, lr, epsilon, state) step(OPT.adagrad, x, dx, lr, epsilon, 0, 0, slot(state, 'm', x)) end   -- OU:33-43
function rmsprop(x, dx, lr, alpha, epsilon, state)                                                       -- OU:46-57
  step(OPT.rmsprop, x, dx, lr, alpha, epsilon, 0, slot(state, 'm', x))
end
function adam(x, dx, lr, beta1, beta2, epsilon, state)                                                   -- OU:59-87
  beta1, beta2, epsilon = beta1 or 0.9, beta2 or 0.999, epsilon or 1e-8
  local m, v = slot(state, 'm', dx), slot(state, 'v', dx)
  state.t = (state.t or 0) + 1
  step(OPT.adam, x, dx, lr, beta1, beta2, epsilon, m, v, state.t)
end
