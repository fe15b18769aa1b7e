set -x
mkdir -p gpurun_out
T=r03f
NCUT="ncu --set full --clock-control none --profile-from-start off -f"
B="python bench.py --steps 1 --no-e2e --no-cpu-baseline --no-also --no-parity"
for r in 2048 3072 4096 8192; do
  CVMX_PLAN_RHI=$r timeout 300 python bench.py --config cfg2 --steps 10 --no-e2e --no-cpu-baseline --no-parity --no-also > gpurun_out/${T}_cfg2_r$r.json 2>/dev/null
  CVMX_PLAN_RHI=$r BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_gram -c 1 -o /tmp/t_r$r $B --config cfg2 > /dev/null 2>&1
  python tools/ncu_summary.py /tmp/t_r$r.ncu-rep | grep -E "duration|grid|DMMA|DRAM read  |DRAM write  |L2 hit" > gpurun_out/${T}_ncu_r$r.txt
  python - <<P
import json
for line in open('gpurun_out/${T}_cfg2_r$r.json'):
    if line.startswith('{'):
        d=json.loads(line); r=d['roofline']; print('rhi=$r', round(d['value'],1), round(d['ms_per_step'],4), 'kernel', round(r['kernel_ms_per_step'],4), r['issued_frac_of_peak'], 'reduce', round(r['reduce_ms_per_step'],3))
P
  cat gpurun_out/${T}_ncu_r$r.txt
done
