# A/B runs of tuning switches (one bench line each): resident and e2e ms per step
run() { echo "$* -> $(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))")"; }
run RAU_SIDE_CTAS=84
run RAU_SIDE_CTAS_FWD=64
run RAU_SIDE_CTAS_FWD=72
run RAU_SIDE_CTAS_FWD=96
run RAU_SIDE_CTAS_BWD=92
run RAU_SIDE_CTAS_BWD=100
run RAU_SIDE_CTAS_BWD=92 RAU_SIDE_CTAS_FWD=72
run RAU_SIDE_CTAS_BWD=100 RAU_SIDE_CTAS_FWD=64
run RAU_SIDE_CTAS=84
