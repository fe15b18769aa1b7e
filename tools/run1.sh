set -x
mkdir -p gpurun_out
T=${1:-r03b}
timeout 900 python -m pytest tests/test_gpu_fused_stats.py -x -q 2>&1 | tail -12 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
for fz in 0 1 0 1; do
  CVMX_FUSE_STATS=$fz timeout 300 python bench.py --config cfg3 --steps 8 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${T}_cfg3_f$fz.json 2> gpurun_out/${T}_cfg3_f$fz.err
  python - <<P
import json
for line in open('gpurun_out/${T}_cfg3_f$fz.json'):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('fuse=$fz', round(d['value']), round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms_per_step'],4), r['issued_frac_of_peak'], 'stats', round(r['stats_ms_per_step'],4))
P
done
timeout 300 python bench.py --config cfg2 --steps 10 --no-e2e --no-cpu-baseline --no-parity --no-also > gpurun_out/${T}_cfg2.json 2>/dev/null
python - <<P
import json
for line in open('gpurun_out/${T}_cfg2.json'):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('cfg2', round(d['value'],1), round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms_per_step'],4), r['issued_frac_of_peak'])
P
