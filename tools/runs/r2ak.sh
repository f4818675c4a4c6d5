#!/bin/bash
# two GPUs, FINAL build: the in-library dispatcher and concurrency tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_concurrency.py tests/test_gpu_multi_device.py "tests/test_gpu_parity.py::test_multi_device_split_matches_single" -m gpu -q > gpurun_out/r2ak_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2ak_pytest_2gpu.log
