-- lua/rau/ffi.lua -- LuaJIT FFI binding of librau.so (include/rau.h).
--
-- The header is plain C and is handed to ffi.cdef verbatim (minus preprocessor lines), so the Lua side and
-- the Python cffi side used by tests/ bind exactly the same declarations.  No arithmetic lives in the Lua
-- shims: they marshal tensors (raw device pointers taken from the CudaTensors AT EVERY CALL, because
-- getParameters()/clone()/share() re-point storages, F:322-347) and turn a non-zero status into error().
--
-- UNEXECUTED IN THIS REPOSITORY'S CI: the build image has no LuaJIT/Torch7 (INTEGRATION.md); the Python
-- mirror rau_vqa_b200/ exercises the same ABI entry points with the same call sequences.
local ffi = require 'ffi'

local M = {}

local function read_header()
  local dir = os.getenv('RAU_HOME') or '.'
  local f = assert(io.open(dir .. '/include/rau.h', 'r'), 'set RAU_HOME to the rau-b200 checkout')
  local out = {}
  for line in f:lines() do
    -- drop preprocessor lines and the extern "C" braces; everything else is plain C declarations
    if not line:match('^%s*#') and not line:match('^extern "C"') and not line:match('^}%s*$') then
      out[#out + 1] = line
    end
  end
  f:close()
  return table.concat(out, '\n')
end

ffi.cdef(read_header())
M.C = ffi.load((os.getenv('RAU_HOME') or '.') .. '/rau_vqa_b200/librau.so')
M.ffi = ffi

-- status -> Lua error (rau_last_error is thread-local text)
function M.check(status)
  if status ~= 0 then
    error('librau: ' .. ffi.string(M.C.rau_last_error()) .. ' (status ' .. tonumber(status) .. ')', 2)
  end
end

-- one native context per device, held in this file-local table: modules never store cdata, so they survive
-- torch.MemoryFile serialisation inside Module:clone() (utils/model_utils.lua MU:18-24)
local contexts = {}
function M.ctx()
  local dev = cutorch.getDevice()
  local c = contexts[dev]
  if c == nil then
    local out = ffi.new('rau_ctx*[1]')
    M.check(M.C.rau_ctx_create(out, dev - 1, nil))   -- NULL = the legacy default stream cutorch computes on
    c = out[0]
    contexts[dev] = c
    M.check(M.C.rau_set_seed(c, torch.initialSeed()))
  end
  return c
end

-- float* of a contiguous (row-contiguous) torch.CudaTensor; nil -> NULL
function M.fptr(t)
  if t == nil then return nil end
  assert(torch.type(t) == 'torch.CudaTensor', 'librau takes torch.CudaTensor only (no CPU path)')
  assert(t:dim() < 2 or t:stride(t:dim()) == 1, 'rows must be contiguous')
  return ffi.cast('float*', t:data())
end

function M.config(opt)
  local c = ffi.new('rau_config')
  c.V = opt.V; c.embed = opt.embed or 200; c.Hq = opt.Hq or 512; c.nlayer = opt.nlayer or 2
  c.C = opt.C or 512; c.S = opt.S or 196; c.M = opt.M or 512; c.A = opt.A or 256; c.H = opt.H or 512
  c.N = opt.N; c.nHop = opt.nHop or 8; c.T = opt.T or 26
  c.p_embed = opt.p_embed or 0.5; c.p_rnn = opt.p_rnn or 0.5; c.p_q = opt.p_q or 0.5
  c.p_x = opt.p_x or 0.5; c.p_m = opt.p_m or 0.5
  return c
end

return M
