-- Stub of utils/vqa_prepro_loader.lua for bench/ref_torch7_cpu.lua: the loader's interface as the experiment scripts use
-- it (load_data F:189; vocab_size F:204, answer_size F:222, seq_len F:339, answer_dict F:907; train_data:next_batch_feat
-- F:434/451, set_batch_order_option / reorder F:711-712, iter_per_epoch F:762; test_data:inorder / iter_per_epoch /
-- next_batch_feat F:853-876) serving synthetic batches of the benchmark's shape.  It also holds the stopwatch: the interval
-- between two consecutive train next_batch_feat calls is one whole iteration of the reference (feval + adam x3 + logging).
local loader = {}

function loader.load_data(vqa_dir, batch_size, prefetch, split, test_batch_size)
  local cfg = assert(RAU_BENCH, 'run through bench/ref_torch7_cpu.lua')
  torch.manualSeed(123)
  local B, T = cfg.batch, cfg.T
  local feats = torch.randn(B, cfg.C, cfg.w, cfg.h):cmax(0)               -- DoubleTensor, post-ReLU statistics (LD:856)
  local x_len = torch.Tensor(B):random(8, T)
  local x = torch.Tensor(T, B):random(2, cfg.V)
  for b = 1, B do for t = x_len[b] + 1, T do x[t][b] = 1 end end            -- ZEROPAD = 1 (LD:1335)
  local y = torch.Tensor(B):random(1, cfg.N)
  local qids = torch.range(1, B)
  local answer_dict = {}
  for i = 1, cfg.N do answer_dict[i] = 'a' .. i end
  local timer, times, calls = torch.Timer(), {}, 0
  local train = {iter_per_epoch = cfg.iters}
  function train:next_batch_feat()
    calls = calls + 1
    if calls > 1 then times[#times + 1] = timer:time().real end          -- the iteration that just ended
    timer:reset()
    return feats, x, x_len, y, qids
  end
  function train:set_batch_order_option() end
  function train:reorder() end
  local test = {iter_per_epoch = 0}
  function test:inorder()                                                   -- called once the training iterations are done
    times[#times + 1] = timer:time().real
    table.remove(times, 1)                                                  -- warm-up iteration
    local s = 0
    for _, t in ipairs(times) do s = s + t end
    local mean = s / #times
    print(string.format('{"impl": "torch7-reference", "metric": "RAU fwd+bwd+update samples/sec", "value": %.4f, ' ..
                        '"unit": "samples/s", "ms_per_step": %.2f, "steps": %d, "batch": %d, "C": %d, "threads": %d}',
                        B / mean, mean * 1e3, #times, B, cfg.C, torch.getnumthreads()))
  end
  function test:next_batch_feat() error('the stub serves no test batches') end
  return {vocab_size = cfg.V, answer_size = cfg.N, seq_len = T, answer_dict = answer_dict, train_data = train, test_data = test}
end

return loader
