#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py cfg3 > gpurun_out/r2p_e2e_probe.log 2>&1; echo "probe rc=$?"; grep -E "wall" gpurun_out/r2p_e2e_probe.log
