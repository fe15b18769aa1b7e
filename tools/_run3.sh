timeout 900 python -m pytest tests/test_gpu_stream_fit.py tests/test_gpu_scan.py -x -q 2>&1 | tail -8
timeout 600 python tools/bench_cfg5.py --rows 200000 --steps 1 > gpurun_out/cfg5_small.json 2> gpurun_out/cfg5_small.err; tail -5 gpurun_out/cfg5_small.err; cat gpurun_out/cfg5_small.json
