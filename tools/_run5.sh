timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused_fit.py -x -q -k "sharded or fused_equals or torch_device" 2>&1 | tail -3
echo "== emulate 8 shards"
BENCH_EMULATE_SHARDS=8 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'], d['roofline']['reduce_ms_per_step'], d['gpu_launches'])"
