set -x
timeout 900 python -m pytest tests/test_gpu_scan.py -x -q 2>&1 | tail -15
timeout 600 python tools/scan_bench.py > gpurun_out/scan_bench.json 2> gpurun_out/scan_bench.err; tail -3 gpurun_out/scan_bench.err; cat gpurun_out/scan_bench.json
