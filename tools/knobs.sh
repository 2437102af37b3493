# A/B runs of tuning switches (one bench line each)
run() { echo "$* -> $(env "$@" python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['ms_per_step'],3), round(d['e2e']['ms_per_step'],3))")"; }
run RAU_WAVE_CAP_A=0
run RAU_WAVE_CAP_A=40 RAU_WAVE_CAP_B=24
run RAU_WAVE_CAP_A=32 RAU_WAVE_CAP_B=32
run RAU_WAVE_CAP_A=48 RAU_WAVE_CAP_B=16
run RAU_WAVE_CAP_A=24 RAU_WAVE_CAP_B=16
