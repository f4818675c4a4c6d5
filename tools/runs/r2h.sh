#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_probe.py cfg3 > gpurun_out/r2h_e2e_probe.log 2>&1; echo "probe rc=$?"; grep -E "wall|chunks" gpurun_out/r2h_e2e_probe.log; grep -A14 "chunk lane" gpurun_out/r2h_e2e_probe.log | head -40; tail -8 gpurun_out/r2h_e2e_probe.log
