set -x
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 500 $T 29591 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02h_bench_cfg2_n$N.json 2> gpurun_out/r02h_bench_cfg2_n$N.err
python - <<P
import json
for line in open('gpurun_out/r02h_bench_cfg2_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_ms'], d['parity']['xtx'], d['parity']['stats_bit_exact'], d['roofline']['traffic'], {k:(round(v['value']),round(v['ms_per_step'],3)) for k,v in d['also'].items()})
P
if [ "$N" = "2" ]; then timeout 300 python -m pytest tests/test_gpu_multi.py tests/test_gpu_slab.py -q 2>&1 | tail -3 > gpurun_out/r02h_pytest_n2.txt; cat gpurun_out/r02h_pytest_n2.txt; fi
grep -i "Traceback" gpurun_out/r02h_bench_cfg2_n$N.err | head -3
