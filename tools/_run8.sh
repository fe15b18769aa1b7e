timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "torch_device" 2>&1 | tail -5
for c in cfg3 cfg4; do
python bench.py --config $c --no-cpu-baseline --steps 5 --e2e-out pinned > gpurun_out/bench_${c}_pinned.json 2> gpurun_out/bench_${c}_pinned.err; tail -2 gpurun_out/bench_${c}_pinned.err; python -c "
import json; d=json.load(open('gpurun_out/bench_${c}_pinned.json')); print('$c', d['value'], d['ms_per_step'], d['e2e'])"
done
