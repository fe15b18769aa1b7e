timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; tail -5 gpurun_out/bench_cfg2_n2.err; cat gpurun_out/bench_cfg2_n2.json
