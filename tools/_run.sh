python -m pytest tests/test_gpu_feed.py -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s3_bench_full.json 2> gpurun_out/s3_bench_full.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s3_bench_full.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e'], d['clocks'], d['roofline']['frac'])
for k,v in d['extra'].items():
    if k!='sweep': print(k, {a:b for a,b in v.items() if a in ('value','ms_per_step','h2d_bytes_per_step','clocks')})
PY
