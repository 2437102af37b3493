# end-of-round evidence on one GPU: full GPU test suite, smoke(), the default bench line, the replayed step's timeline,
# the launch list of four eager steps (ncu gpu__time_duration, cold caches)
set -x
O=gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > $O/r02d_gputests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $O/r02d_smoke.txt 2>&1
python bench.py > $O/r02d_bench_ours_full_mixed.json 2> $O/r02d_bench.err
RAU_PHASES=2 python tools/phases.py ours_full > $O/r02d_timeline.txt 2>&1
RAU_GRAPH=0 ncu --clock-control none --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r02d_launches_ours_full.csv python tools/phases.py ours_full > $O/r02d_launches.log 2>&1
cat $O/r02d_gputests.txt; tail -2 $O/r02d_smoke.txt; head -c 400 $O/r02d_bench_ours_full_mixed.json
