for s in 0 1; do
  echo "== emulate 8 shards, CVMX_SCAN=$s"
  CVMX_SCAN=$s BENCH_EMULATE_SHARDS=8 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'], d['roofline']['reduce_ms_per_step'], d['gpu_launches'])"
done
echo "== 1 GPU default"
python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'], d['gpu_launches'])"
