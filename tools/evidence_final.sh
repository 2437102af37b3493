# end-of-round evidence on one GPU (round 2, last code): full GPU test suite, smoke(), the default bench line, the replayed
# step's timeline, the launch list of four eager steps (ncu gpu__time_duration, cold caches), ncu --set full of the roofline
# kernel (i_embed product, default mode, 148 SMs) and of the encoder's hoisted input projection (CTA pairs)
set -x
O=gpurun_out
T=r02e
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/${T}_gputests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/${T}_smoke.txt 2>&1
python bench.py > $O/${T}_bench_ours_full_mixed.json 2> $O/${T}_bench.err
RAU_PHASES=2 python tools/phases.py ours_full > $O/${T}_timeline.txt 2>&1
RAU_ROWS_TRACE=1 python tools/iembed_trace.py mixed > $O/${T}_iembed_trace.txt 2>&1
RAU_GRAPH=0 ncu --clock-control none --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/${T}_launches_ours_full.csv python tools/phases.py ours_full > $O/${T}_launches.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:rows_gemm_kernel -c 2 -o $O/${T}_tanh_mixed python tools/ncu_iembed.py mixed > $O/${T}_tanh.log 2>&1
ncu --clock-control none --set full --import-source on -k regex:rows_gemm_kernel -s 2 -c 2 -o $O/${T}_enc_proj python tools/proj_trace.py > $O/${T}_enc_proj.log 2>&1
cat $O/${T}_gputests.txt; tail -2 $O/${T}_smoke.txt; head -c 400 $O/${T}_bench_ours_full_mixed.json
