-- bench/ref_torch7_cpu.lua -- times the REFERENCE's own training iteration (feval + the three adam calls, the span its
-- torch.Timer measures at F:786-795 and prints as `time=` at F:848-849) on CPU, for anyone who has a Torch7 install.
-- It cannot run in the build image (no LuaJIT / torch rocks, SURVEY.md 8c) and is shipped unexecuted.
--
--   th bench/ref_torch7_cpu.lua -ref /path/to/RAU_VQA [-variant Ours_SS] [-batch 8] [-iters 5] [-C 512] [-nhop 1]
--
-- How: the unmodified experiment script is run with `-gpuid -1 -backend nn` (only Ours_SS and Ours_ResNet have a CPU path,
-- SURVEY.md 0.3) while `require 'utils.vqa_prepro_loader'` resolves to the stub next to this file, which hands the script
-- synthetic batches of the benchmark's shape (SURVEY.md 8d) through the loader's own return tuple
-- (feats[B,C,w,h], x[T,B], x_len[B], y[B], qids; LD:1009) and clocks the interval between two next_batch_feat calls --
-- one whole iteration.  Nothing of the reference is copied: its model, feval and optimizer run as they are.
local cmd = torch.CmdLine()
cmd:option('-ref', '', 'root of a RAU_VQA checkout')
cmd:option('-variant', 'Ours_SS', 'Ours_SS | Ours_ResNet (the variants with a CPU path)')
cmd:option('-batch', 8, 'batch size (BASELINE.json configs[0]: 8)')
cmd:option('-iters', 5, 'timed iterations (after 1 warm-up)')
cmd:option('-C', 512, 'feature channels: 512 (VGG16 pool5) or 2048 (ResNet-101)')
cmd:option('-nhop', 1, 'answering units')
local o = cmd:parse(arg)
assert(o.ref ~= '', 'give -ref /path/to/RAU_VQA')

local here = debug.getinfo(1, 'S').source:match('^@(.*)/[^/]*$') or '.'
-- 1. the stub loader wins over the reference's; 2. model/, the other utils/ come from the reference tree
package.path = here .. '/torch7/?.lua;' .. o.ref .. '/?.lua;' .. package.path
RAU_BENCH = {batch = o.batch, iters = o.iters + 1, C = o.C, T = 26, V = 16384, N = 2000, w = 14, h = 14}

local script = o.ref .. '/experiments/' .. o.variant .. '/LstmAttCtrlGradNoiseDontSelect.lua'
arg = {'-gpuid', '-1', '-backend', 'nn', '-display', 'false', '-visatt', 'false', '-max_epochs', '1', '-split', 'test-dev2015',
       '-batch_size', tostring(o.batch), '-test_batch_size', tostring(o.batch), '-nhop', tostring(o.nhop),
       '-cnnout_w', '14', '-cnnout_h', '14', '-save_dir', os.tmpname() .. '_rau_bench'}
if o.variant ~= 'Ours_SS' then arg[#arg + 1] = '-cnnout_dim'; arg[#arg + 1] = tostring(o.C) end
dofile(script)     -- runs RAU_BENCH.iters iterations, then the (empty) test pass; the stub prints the result line
