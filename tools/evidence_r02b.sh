set -x
O=gpurun_out
NCU="ncu --clock-control none"
python bench.py > $O/r02b_bench_ours_full_mixed.json 2> $O/r02b_bench.err
RAU_PHASES=2 python tools/phases.py ours_full > $O/r02b_timeline.txt 2>&1
RAU_GRAPH=0 $NCU --metrics gpu__time_duration.sum -c 2500 --csv --log-file $O/r02b_launches_ours_full.csv python tools/phases.py ours_full > $O/r02b_launches.log 2>&1
RAU_GRAPH=0 $NCU --set full --import-source on --kernel-name-base demangled -k regex:lstm_seq_kernel -c 2 -o $O/r02b_lstm_seq python tools/phases.py ours_full > $O/r02b_lstm_seq.log 2>&1
$NCU --set full -k 'regex:attn_rows|xprep_rows' -o $O/r02b_hbm_kernels python tools/ncu_sweep.py mixed > $O/r02b_hbm_kernels.log 2>&1
$NCU --set full --import-source on -k regex:rows_gemm_kernel -s 4 -c 2 -o $O/r02b_dy python tools/ncu_sweep.py mixed > $O/r02b_dy.log 2>&1
ls -la $O/*.ncu-rep
tail -c 600 $O/r02b_bench_ours_full_mixed.json
