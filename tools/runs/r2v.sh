#!/bin/bash
mkdir -p gpurun_out
date +%T
timeout 1100 python -m pytest tests -m gpu -q -x --durations=12 > gpurun_out/r2v_pytest_gpu.log 2>&1
echo "pytest rc=$?"; date +%T; tail -22 gpurun_out/r2v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1
echo "smoke rc=$?"; date +%T; tail -3 gpurun_out/r2v_smoke.log
timeout 600 python bench.py > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err
echo "bench rc=$?"; date +%T; cut -c1-1500 gpurun_out/r2v_bench.json
