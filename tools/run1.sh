set -x
mkdir -p gpurun_out
T=${1:-r02w}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
timeout 500 python bench.py > gpurun_out/${T}_bench_cfg2.json 2> gpurun_out/${T}_bench_cfg2.err
CVMX_HOST_STAGER=0 timeout 500 python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_cfg2_nostager.json 2> gpurun_out/${T}_bench_cfg2_nostager.err
for t in 2 4 8 16; do CVMX_STAGE_THREADS=$t timeout 300 python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench_cfg2_t$t.json 2>/dev/null; done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_bench_cfg2*.json')):
  for line in open(f):
    if line.startswith('{'):
        d=json.loads(line); print(f, round(d['value'],1), round(d['e2e']['value'],2), d['e2e']['breakdown_ms']['partitioner_ms'], d['e2e'].get('pageable_input'))
P
timeout 900 python tools/parity_report.py > gpurun_out/${T}_parity.log 2>&1; tail -3 gpurun_out/${T}_parity.log
