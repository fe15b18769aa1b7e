#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU test suite, smoke, benches, ncu launch lists and captures.
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/final_pytest_gpu.txt; cat gpurun_out/final_pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 > gpurun_out/final_smoke.txt; cat gpurun_out/final_smoke.txt
timeout 400 python bench.py > gpurun_out/final_bench_cfg2.json 2> gpurun_out/final_bench_cfg2.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
timeout 300 python bench.py --config cfg3 --no-cpu-baseline --steps 5 > gpurun_out/final_bench_cfg3.json 2> gpurun_out/final_bench_cfg3.err
timeout 300 python bench.py --config cfg4 --no-cpu-baseline --steps 5 > gpurun_out/final_bench_cfg4.json 2> gpurun_out/final_bench_cfg4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg2.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
BENCH_EMULATE_SHARDS=8 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_cfg2_shard8.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:"k_scan_segsums|k_scan_delta|k_scan_chain|k_scan_prefix" -c 4 -o gpurun_out/r01_scan_full python tools/scan_once.py > /dev/null 2>&1
ls -la gpurun_out | tail -20
