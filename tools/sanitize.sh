#!/bin/bash
# compute-sanitizer over the small configurations: memcheck on every case, racecheck on the
# kernels that hand-roll shared-memory protocols (mbarrier rings, setmaxnreg warp roles, the reused epilogue tile),
# initcheck on the README shape.  Run on the GPU box:  bash tools/sanitize.sh  ->  gpurun_out/sanitizer_*.log
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck; do
  cases="all"; [ "$tool" = racecheck ] && cases="cfg1 split scan loo"
  for c in $cases; do
    timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitizer_case.py $c \
        > gpurun_out/sanitizer_${tool}_${c}.log 2>&1
    echo "$tool $c exit=$?" | tee -a gpurun_out/sanitizer_summary.txt
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|SANITIZER_CASES_OK" gpurun_out/sanitizer_${tool}_${c}.log | tee -a gpurun_out/sanitizer_summary.txt
  done
done
timeout 600 $CS --tool initcheck --print-limit 20 --error-exitcode 9 python tools/sanitizer_case.py cfg1 > gpurun_out/sanitizer_initcheck_cfg1.log 2>&1
echo "initcheck cfg1 exit=$?" | tee -a gpurun_out/sanitizer_summary.txt
grep -E "ERROR SUMMARY|SANITIZER_CASES_OK" gpurun_out/sanitizer_initcheck_cfg1.log | tee -a gpurun_out/sanitizer_summary.txt
