# A/B runs of the SM split between the chain and the side stream (one bench line each)
run() { echo "$* -> $(env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))")"; }
run RAU_SIDE_CTAS=84
run RAU_SIDE_CTAS_BWD=96
run RAU_SIDE_CTAS_BWD=108
run RAU_SIDE_CTAS_BWD=120
run RAU_SIDE_CTAS=76 RAU_SIDE_CTAS_BWD=100
run RAU_SIDE_CTAS=76
