timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/dist_fit_worker.py 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -15
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_cfg2_n2.json 2> gpurun_out/bench_cfg2_n2.err; tail -3 gpurun_out/bench_cfg2_n2.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg2_n2.json')); print(d['value'], d['ms_per_step'], d['config']['parallelism'])"
CVMX_PEER_REDUCE=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29546 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['parallelism'])"
