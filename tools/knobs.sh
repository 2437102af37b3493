for s in 40 56 64 72 84 100; do
  echo "SIDE=$s $(RAU_SIDE_CTAS=$s python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['e2e']['ms_per_step'])")"
done
for f in 40 64; do for b in 56 72; do
  echo "FWD=$f BWD=$b $(RAU_SIDE_CTAS_FWD=$f RAU_SIDE_CTAS_BWD=$b python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'])")"
done; done
