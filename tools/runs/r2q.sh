#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2q_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2q_pytest_gpu.log
timeout 600 python tools/e2e_probe.py cfg3 > gpurun_out/r2q_e2e_probe.log 2>&1; echo "probe rc=$?"; grep -E "wall" gpurun_out/r2q_e2e_probe.log
