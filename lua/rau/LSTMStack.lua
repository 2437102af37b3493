-- lua/rau/LSTMStack.lua -- nn.RauLSTMStack: the nn.Module behind both LSTM factories (model/ATTLSTM.lua,
-- model/DeepLSTM.lua).  One rau_lstm_cell_fwd / rau_lstm_cell_bwd call per layer, rau_dropout between layers.
-- Parameters sit in child nn.Linear modules listed in self.modules, and the class derives from nn.Container: stock
-- nn.Module:share only touches self[name], it is nn.Container:share that recurses into self.modules -- so
-- proto:clone('weight','bias','gradWeight','gradBias') (F:339-347) and nngraph's recursion into the attlstm node (F:273)
-- really alias the clones' weights and gradients to the prototype's.  parameters(), getParameters(), training(),
-- evaluate(), zeroGradParameters() are nn.Container's own (SURVEY 8b).
-- Python mirror with identical call sequence: rau_vqa_b200/model/_lstm_stack.py.
require 'nn'
require 'cutorch'
local rau = require 'rau.ffi'
local C, ffi = rau.C, rau.ffi

local Stack, parent = torch.class('nn.RauLSTMStack', 'nn.Container')

local GATES_IFOG, GATES_IGFO = 0, 1
local stream_counter = 0

-- packed_state: DeepLSTM keeps [c1|h1|c2|h2] in one tensor (D:23-24); ATTLSTM takes {x, c, h} (A:30)
function Stack:__init(input_size, rnn_size, num_layers, dropout, gate_order, packed_state, dropout_on_first)
  parent.__init(self)
  self.input_size, self.rnn_size, self.num_layers = input_size, rnn_size, num_layers
  self.dropout = dropout or 0
  self.gate_order, self.packed_state, self.dropout_on_first = gate_order, packed_state, dropout_on_first
  for L = 1, num_layers do   -- (nn.Container:add appends to self.modules)
    self:add(nn.Linear(L == 1 and input_size or rnn_size, 4 * rnn_size))   -- i2h (A:6 / D:43)
    self:add(nn.Linear(rnn_size, 4 * rnn_size))                            -- h2h (A:7 / D:44)
  end
  self.train = true
  self.stream_id = 0
  self.saved = {}      -- per layer {input after dropout, 5*B*H saved gate activations}; private to each clone
  self.gradInput = {}
end

local function views(self, state, L)
  local H = self.rnn_size
  if self.packed_state then
    return state:narrow(2, 2 * (L - 1) * H + 1, H), state:narrow(2, 2 * (L - 1) * H + H + 1, H)
  end
  return state[1]:narrow(2, (L - 1) * H + 1, H), state[2]:narrow(2, (L - 1) * H + 1, H)
end

local function has_dropout(self, L) return self.dropout > 0 and (L > 1 or self.dropout_on_first) end

local function desc(self, B, in_size, x, c_prev, h_prev, c, h)
  local d = ffi.new('rau_lstm_desc')
  d.B, d.in_size, d.H, d.gate_order = B, in_size, self.rnn_size, self.gate_order
  d.ldx, d.ldc_prev, d.ldh_prev, d.ldc, d.ldh = x:stride(1), c_prev:stride(1), h_prev:stride(1), c:stride(1), h:stride(1)
  return d
end

local function split_input(self, input)
  if self.packed_state then return input[1], input[2] end
  return input[1], {input[2], input[3]}
end

function Stack:updateOutput(input)
  local x, state = split_input(self, input)
  local B, H, n = x:size(1), self.rnn_size, self.num_layers
  local ctx = rau.ctx()
  stream_counter = stream_counter + 1
  self.stream_id = stream_counter
  local out_state
  if self.packed_state then
    self.output = (torch.type(self.output) == 'torch.CudaTensor') and self.output or torch.CudaTensor()
    self.output:resize(B, 2 * n * H)
    out_state = self.output
  else
    if torch.type(self.output) ~= 'table' then self.output = {torch.CudaTensor(), torch.CudaTensor()} end
    self.output[1]:resize(B, n * H); self.output[2]:resize(B, n * H)
    out_state = self.output
  end
  local u = x
  for L = 1, n do
    local in_size = (L == 1) and self.input_size or H
    local sv = self.saved[L] or {torch.CudaTensor(), torch.CudaTensor()}
    self.saved[L] = sv
    sv[1]:resize(B, in_size); sv[2]:resize(5, B, H)
    if has_dropout(self, L) then
      local uc = u:isContiguous() and u or u:contiguous()
      rau.check(C.rau_dropout(ctx, uc:nElement(), rau.fptr(uc), self.dropout, self.train and 1 or 0, nil,
                              self.stream_id * 64 + (L - 1), rau.fptr(sv[1])))
    else
      sv[1]:copy(u)
    end
    local c_prev, h_prev = views(self, state, L)
    local c_new, h_new = views(self, out_state, L)
    local i2h, h2h = self.modules[2 * L - 1], self.modules[2 * L]
    rau.check(C.rau_lstm_cell_fwd(ctx, desc(self, B, in_size, sv[1], c_prev, h_prev, c_new, h_new),
                                  rau.fptr(sv[1]), rau.fptr(c_prev), rau.fptr(h_prev),
                                  rau.fptr(i2h.weight), rau.fptr(i2h.bias), rau.fptr(h2h.weight), rau.fptr(h2h.bias),
                                  rau.fptr(c_new), rau.fptr(h_new), rau.fptr(sv[2])))
    u = h_new
  end
  return self.output
end

-- updateGradInput and accGradParameters share one native call per layer; want_params == false passes NULL
-- gradient pointers (the call then only forms gradInput), want_input == false discards gradInput.
local function bwd(self, input, gradOutput, scale, want_input, want_params)
  local x, state = split_input(self, input)
  local B, H, n = x:size(1), self.rnn_size, self.num_layers
  local ctx = rau.ctx()
  local g_state, d_state
  if self.packed_state then
    g_state = gradOutput
    d_state = torch.CudaTensor(B, 2 * n * H):zero()
  else
    g_state = {gradOutput[1], gradOutput[2]}
    d_state = {torch.CudaTensor(B, n * H):zero(), torch.CudaTensor(B, n * H):zero()}
  end
  local from_above, dx = nil, nil
  for L = n, 1, -1 do
    local in_size = (L == 1) and self.input_size or H
    local sv = self.saved[L]
    local c_prev, h_prev = views(self, state, L)
    local dc_out, dh_out = views(self, g_state, L)
    local dc_prev, dh_prev = views(self, d_state, L)
    local du = torch.CudaTensor(B, in_size)
    local i2h, h2h = self.modules[2 * L - 1], self.modules[2 * L]
    local gp = function(t) if want_params then return rau.fptr(t) end return nil end
    local d = desc(self, B, in_size, sv[1], c_prev, h_prev, dc_prev, dh_prev)
    d.ldc, d.ldh = H, H
    rau.check(C.rau_lstm_cell_bwd(ctx, d, rau.fptr(sv[1]), rau.fptr(c_prev), rau.fptr(h_prev),
                                  rau.fptr(i2h.weight), rau.fptr(h2h.weight), rau.fptr(sv[2]),
                                  rau.fptr(dc_out), rau.fptr(dh_out), dc_out:stride(1), dh_out:stride(1), rau.fptr(from_above),
                                  rau.fptr(du), rau.fptr(dc_prev), rau.fptr(dh_prev), du:stride(1), dc_prev:stride(1), dh_prev:stride(1),
                                  gp(i2h.gradWeight), gp(i2h.gradBias), gp(h2h.gradWeight), gp(h2h.gradBias), scale))
    if has_dropout(self, L) then   -- nn.Dropout backward = the same mask applied to the gradient
      local dud = torch.CudaTensor(B, in_size)
      rau.check(C.rau_dropout(ctx, du:nElement(), rau.fptr(du), self.dropout, self.train and 1 or 0, nil,
                              self.stream_id * 64 + (L - 1), rau.fptr(dud)))
      du = dud
    end
    if L > 1 then from_above = du else dx = du end
  end
  if want_input then
    if self.packed_state then self.gradInput = {dx, d_state} else self.gradInput = {dx, d_state[1], d_state[2]} end
  end
  return self.gradInput
end

-- A container that drives its children with updateGradInput and accGradParameters as two calls (instead of
-- :backward) must not pay for two native backward passes: updateGradInput runs the ONE native call with the weight
-- gradients directed into a zeroed scratch set (file-local, shared by all clones: the pair of calls for one module is
-- never interleaved with another module's pair inside one container pass), and the matching accGradParameters only adds
-- scale * scratch into the shared gradWeight / gradBias.  Any other call order falls back to the native call.
local pending = {owner = nil, inp = nil, gout = nil, scratch = {}}

local function ptr_of(t)
  if torch.type(t) == 'table' then return ptr_of(t[1]) end
  return tonumber(ffi.cast('intptr_t', rau.fptr(t)))
end

local function scratch_like(key, t)
  local s = pending.scratch[key]
  if s == nil or s:nElement() ~= t:nElement() then s = t.new():resizeAs(t); pending.scratch[key] = s end
  return s:zero()
end

function Stack:updateGradInput(input, gradOutput)
  local real = {}
  for i, m in ipairs(self.modules) do      -- divert the weight gradients of this one call into the scratch set
    real[i] = {m.gradWeight, m.gradBias}
    m.gradWeight, m.gradBias = scratch_like(2 * i - 1, m.gradWeight), scratch_like(2 * i, m.gradBias)
  end
  local ok, err = pcall(bwd, self, input, gradOutput, 1, true, true)
  for i, m in ipairs(self.modules) do m.gradWeight, m.gradBias = real[i][1], real[i][2] end
  if not ok then error(err) end
  pending.owner, pending.inp, pending.gout = self, ptr_of(input), ptr_of(gradOutput)
  return self.gradInput
end

function Stack:accGradParameters(input, gradOutput, scale)
  scale = scale or 1
  if pending.owner == self and pending.inp == ptr_of(input) and pending.gout == ptr_of(gradOutput) then
    for i, m in ipairs(self.modules) do
      m.gradWeight:add(scale, pending.scratch[2 * i - 1])
      m.gradBias:add(scale, pending.scratch[2 * i])
    end
    pending.owner = nil
    return
  end
  bwd(self, input, gradOutput, scale, false, true)
end

function Stack:backward(input, gradOutput, scale)
  pending.owner = nil
  return bwd(self, input, gradOutput, scale or 1, true, true)
end

return {Stack = Stack, GATES_IFOG = GATES_IFOG, GATES_IGFO = GATES_IGFO}
