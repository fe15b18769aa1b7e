timeout 900 python -m pytest tests/test_gpu_fused_fit.py tests/test_gpu_scan.py tests/test_gpu_stream_fit.py -x -q 2>&1 | tail -6
python bench.py --no-cpu-baseline --steps 10 > gpurun_out/bench_cfg2_fused.json 2> gpurun_out/bench_cfg2_fused.err; tail -3 gpurun_out/bench_cfg2_fused.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_fused.json')); print(d['value'], d['ms_per_step'], d['e2e']); print(d['roofline'])"
