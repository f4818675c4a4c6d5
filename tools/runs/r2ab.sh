#!/bin/bash
# FINAL build: full GPU suite, smoke, default bench, ncu launch list of the same bench command
mkdir -p gpurun_out
date +%T
timeout 1100 python -m pytest tests -m gpu -q --durations=6 > gpurun_out/r2ab_pytest_gpu.log 2>&1
echo "pytest rc=$?"; date +%T; tail -12 gpurun_out/r2ab_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2ab_smoke.log
timeout 600 python bench.py > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err
echo "bench rc=$?"; date +%T; cut -c1-330 gpurun_out/r2ab_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2ab_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ab_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"; date +%T; grep -c "pack_seed_final_kernel" gpurun_out/r2ab_launches.csv; grep "pack_seed_final_kernel\|count_kmers_oct_kernel" gpurun_out/r2ab_launches.csv | tail -4 | cut -c1-260
