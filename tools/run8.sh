set -x
mkdir -p gpurun_out
N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 500 $T 29551 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02v_bench_cfg2_n$N.json 2> gpurun_out/r02v_bench_cfg2_n$N.err
CVMX_SLAB_TIMING=1 timeout 500 $T 29552 bench.py --gpus $N --steps 5 --warmup 3 --no-also --no-parity > gpurun_out/r02v_timing_n$N.json 2> gpurun_out/r02v_timing_n$N.err
python - <<P
import json
for f in ['gpurun_out/r02v_bench_cfg2_n$N.json','gpurun_out/r02v_timing_n$N.json']:
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line); print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_ms'], d['e2e']['includes'][:260], d['parity'] and d['parity']['xtx'])
P
grep "slab fit rank" gpurun_out/r02v_timing_n$N.err | tail -8
