#!/bin/bash
# round-2 second call: the whole GPU suite on the rewritten host pipeline + staged index build, then the 1.51 Gsymbol
# index with the final-step image under a build trace (round r2a's run of it died silently)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2b_smi.txt 2>&1
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2b_pytest_gpu.log
for fin in 0 1; do
  MSBWT_TRACE=1 MSBWT_FINAL_INDEX=$fin timeout 600 python -X faulthandler tools/pack_ab.py --workload cfg3 > gpurun_out/r2b_cfg3_fin$fin.json 2> gpurun_out/r2b_cfg3_fin$fin.err
  echo "cfg3 fin=$fin rc=$?"; cat gpurun_out/r2b_cfg3_fin$fin.json; tail -25 gpurun_out/r2b_cfg3_fin$fin.err
done
dmesg 2>/dev/null | tail -5
