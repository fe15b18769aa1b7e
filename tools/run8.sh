set -x
mkdir -p gpurun_out
N=${1:-4}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
timeout 500 $T 29581 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02z_bench_cfg2_n$N.json 2> gpurun_out/r02z_bench_cfg2_n$N.err
python - <<P
import json
for line in open('gpurun_out/r02z_bench_cfg2_n$N.json'):
    if line.startswith('{'):
        d=json.loads(line); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['breakdown_ms'], d['parity'], {k:(v['value']) for k,v in d['also'].items()})
P
grep -i "error\|Traceback" gpurun_out/r02z_bench_cfg2_n$N.err | head -5
