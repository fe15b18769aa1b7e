set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_slab.py tests/test_gpu_multi.py tests/test_partitioner.py -x -q 2>&1 | tail -15 > gpurun_out/r02u_pytest.txt; cat gpurun_out/r02u_pytest.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
CVMX_SLAB_TIMING=1 timeout 500 $T 29561 bench.py --gpus 2 --steps 10 --warmup 3 --no-also > gpurun_out/r02u_bench_cfg2_n2.json 2> gpurun_out/r02u_bench_cfg2_n2.err
python - <<P
import json
for line in open('gpurun_out/r02u_bench_cfg2_n2.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_ms'], d['e2e']['includes'][:200], d['parity'])
P
grep "slab fit rank" gpurun_out/r02u_bench_cfg2_n2.err | tail -4
tail -5 gpurun_out/r02u_bench_cfg2_n2.err
