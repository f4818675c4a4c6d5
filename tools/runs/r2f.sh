#!/bin/bash
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2f_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2f_pytest_gpu.log
for w in cfg3 cfg2; do
  timeout 600 python tools/pack_ab.py --workload $w > gpurun_out/r2f_${w}.json 2> gpurun_out/r2f_${w}.err
  echo "$w rc=$?"; cat gpurun_out/r2f_${w}.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2f_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2f_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2f_ncu.log
