set -x
mkdir -p gpurun_out
T=${1:-r02x}
timeout 600 python -m pytest tests/test_gpu_host_stager.py -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest.txt; cat gpurun_out/${T}_pytest.txt
for s in 0 1 2 0 1 2; do
  CVMX_LOO_STORE=$s timeout 300 python bench.py --config cfg4 --steps 8 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${T}_cfg4_s$s.json 2>/dev/null
  python - <<P
import json
for line in open('gpurun_out/${T}_cfg4_s$s.json'):
    if line.startswith('{'):
        d=json.loads(line); print('store hint $s', round(d['value']), round(d['ms_per_step'],4), round(d['roofline']['kernel_ms_per_step'],4), round(d['roofline']['write_gbs']))
P
done
