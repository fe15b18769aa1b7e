timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "golden or midsize or summation or cfg2" 2>&1 | tail -3
timeout 200 python bench.py --no-cpu-baseline --no-e2e --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['stats_ms_per_step'])"
