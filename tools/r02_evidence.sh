#!/bin/bash
# Round-2 evidence on one B200 (run under gpurun): GPU test suite, smoke, compute-sanitizer memcheck over the small
# cases (float32 tcgen05 kernel included), ncu --set full captures of the fold-path kernels (tools/prof_once.py),
# launch lists, the default bench line and the reference arm.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-r02m}
set -x
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/${TAG}_pytest.txt; cat gpurun_out/${TAG}_pytest.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/${TAG}_smoke.txt; cat gpurun_out/${TAG}_smoke.txt
for c in all f32; do
  timeout 900 $CS --tool memcheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_case.py $c > gpurun_out/${TAG}_memcheck_$c.log 2>&1
  echo "memcheck $c exit=$?" | tee -a gpurun_out/${TAG}_sanitizer_summary.txt
  grep -E "ERROR SUMMARY|SANITIZER_CASES_OK" gpurun_out/${TAG}_memcheck_$c.log | tee -a gpurun_out/${TAG}_sanitizer_summary.txt
done
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off -f"
timeout 300 $NCU -k regex:k_gram -c 1 -o gpurun_out/${TAG}_gram_lmo python tools/prof_once.py lmo > gpurun_out/${TAG}_ncu.log 2>&1
timeout 300 $NCU -k regex:k_gram -c 2 -o gpurun_out/${TAG}_gram_kfold python tools/prof_once.py kfold >> gpurun_out/${TAG}_ncu.log 2>&1
timeout 300 $NCU -k regex:k_loo -c 2 -o gpurun_out/${TAG}_loo python tools/prof_once.py loo >> gpurun_out/${TAG}_ncu.log 2>&1
timeout 300 $NCU -k regex:k_gram_tc -c 1 -o gpurun_out/${TAG}_gram_tc python tools/prof_once.py f32 >> gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log
timeout 500 python bench.py > gpurun_out/${TAG}_bench_cfg2.json 2> gpurun_out/${TAG}_bench_cfg2.err; tail -c 600 gpurun_out/${TAG}_bench_cfg2.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2> gpurun_out/${TAG}_bench_ref.err; cat gpurun_out/${TAG}_bench_ref.json
for c in cfg2 cfg3 cfg4; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches_$c.csv python bench.py --config $c --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-also > /dev/null 2>&1
done
# DRAM traffic of the dominant kernel per step (roofline.traffic): ncu --set full over the timed steps only
NCUT="ncu --set full --clock-control none --profile-from-start off -f"
B="python bench.py --steps 1 --no-e2e --no-cpu-baseline --no-also --no-parity"
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_gram -c 1 -o gpurun_out/${TAG}_traffic_cfg2 $B --config cfg2 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_gram -c 1 -o gpurun_out/${TAG}_traffic_cfg3 $B --config cfg3 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 timeout 400 $NCUT -k regex:^k_loo -c 10 -o gpurun_out/${TAG}_traffic_cfg4 $B --config cfg4 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 BENCH_EMULATE_SHARDS=8 timeout 400 $NCUT -k regex:^k_gram -c 1 -o gpurun_out/${TAG}_traffic_cfg2_s8 $B --config cfg2 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 BENCH_EMULATE_SHARDS=2 timeout 400 $NCUT -k regex:^k_gram -c 1 -o gpurun_out/${TAG}_traffic_cfg2_s2 $B --config cfg2 > /dev/null 2>&1
BENCH_CUDA_PROFILER=1 BENCH_EMULATE_SHARDS=4 timeout 400 $NCUT -k regex:^k_gram -c 1 -o gpurun_out/${TAG}_traffic_cfg2_s4 $B --config cfg2 > /dev/null 2>&1
ls -la gpurun_out | tail -30
