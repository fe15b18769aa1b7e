set -x
mkdir -p gpurun_out
T=r03g
for st in 2 3 4 8; do
  CVMX_MOM_STAGES=$st timeout 300 python bench.py --config cfg2 --steps 10 --no-e2e --no-cpu-baseline --no-parity --no-also > gpurun_out/${T}_cfg2_st$st.json 2>/dev/null
  python - <<P
import json
for line in open('gpurun_out/${T}_cfg2_st$st.json'):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('mom stages=$st', round(d['value'],1), round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms_per_step'],4), r['issued_frac_of_peak'], 'stats', round(r['stats_ms_per_step'],3))
P
done
