#!/bin/bash
# round-2 first call: everything round 1 wrote but never ran (multi-device tests skip on a 1-GPU box)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_parity.py::test_multi_device_split_matches_single -m gpu -q > gpurun_out/r2a_multi.log 2>&1
echo "multi rc=$?"; tail -15 gpurun_out/r2a_multi.log
bash tools/gpu_round.sh r2a exp
