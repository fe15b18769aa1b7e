#!/bin/bash
# Round-2 final numbers on one B200 (under gpurun; everything it leaves in gpurun_out/ stays small: ncu reports are exported
# to text on the box and deleted): ncu capture of the fused-statistics Gram kernel, default bench line, reference arm,
# launch lists, DRAM traffic of the dominant kernel per config (profiles/traffic.json).
TAG=${1:-r02g}
set -x
mkdir -p gpurun_out
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off -f"
timeout 300 $NCU -k regex:k_gram -c 1 -o /tmp/${TAG}_gram_lmo python tools/prof_once.py lmo > gpurun_out/${TAG}_ncu.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_gram_lmo.ncu-rep > gpurun_out/${TAG}_gram_lmo_ncu.txt
ncu -i /tmp/${TAG}_gram_lmo.ncu-rep --page raw --csv > gpurun_out/${TAG}_gram_lmo_ncu_raw.csv 2>/dev/null
timeout 500 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err
for c in cfg2 cfg3 cfg4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > /dev/null 2>&1
done
BENCH_EMULATE_SHARDS=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_cfg2_shard8.csv python bench.py --config cfg2 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > /dev/null 2>&1
NCUT="ncu --set full --clock-control none --profile-from-start off -f"
B="python bench.py --steps 1 --no-e2e --no-cpu-baseline --no-also --no-parity"
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_gram -c 1 -o /tmp/t_cfg2 $B --config cfg2 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_gram -c 1 -o /tmp/t_cfg3 $B --config cfg3 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_loo -c 10 -o /tmp/t_cfg4 $B --config cfg4 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 BENCH_EMULATE_SHARDS=8 timeout 400 $NCUT -k regex:^k_gram -c 1 -o /tmp/t_cfg2_s8 $B --config cfg2 > /dev/null 2>&1
python tools/traffic_from_ncu.py cfg2=/tmp/t_cfg2.ncu-rep cfg3=/tmp/t_cfg3.ncu-rep cfg4=/tmp/t_cfg4.ncu-rep cfg2@8=/tmp/t_cfg2_s8.ncu-rep
cp profiles/traffic.json gpurun_out/${TAG}_traffic.json
python tools/ncu_summary.py /tmp/t_cfg3.ncu-rep /tmp/t_cfg2.ncu-rep /tmp/t_cfg2_s8.ncu-rep > gpurun_out/${TAG}_gram_fullsize_ncu.txt
du -sh gpurun_out; ls -la gpurun_out | grep ${TAG}
