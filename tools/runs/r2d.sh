#!/bin/bash
# round-2 call d: the reordered one-request kernel: timing + one ncu --set full capture
mkdir -p gpurun_out
for w in cfg3 cfg2; do
  timeout 600 python tools/pack_ab.py --workload $w > gpurun_out/r2d_${w}.json 2> gpurun_out/r2d_${w}.err
  echo "$w rc=$?"; cat gpurun_out/r2d_${w}.json
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2d_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2d_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2d_ncu.log; ls -la gpurun_out/*.ncu-rep
