python -m pytest tests -m gpu -x -q 2>&1 | tail -4
run() { echo "$* -> $(env "$@" python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))")"; }
run RAU_X=1
run RAU_X=2
run RAU_X=3
RAU_PHASES=2 python tools/phases.py ours_full > gpurun_out/s3_timeline5.txt 2>&1; tail -58 gpurun_out/s3_timeline5.txt | head -14
