timeout 900 python -m pytest tests/test_gpu_scan.py tests/test_gpu_stream_fit.py -x -q 2>&1 | tail -3
echo "== emulate 8 shards"
BENCH_EMULATE_SHARDS=8 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'], d['roofline']['reduce_ms_per_step'], d['gpu_launches'])"
echo "== emulate 4 shards"
BENCH_EMULATE_SHARDS=4 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'], d['roofline']['reduce_ms_per_step'], d['gpu_launches'])"
