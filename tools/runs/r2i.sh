#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_final_fast.py tests/test_gpu_u64_kmers.py tests/test_gpu_host_pack.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2i_pytest.log
for deep in 1 0; do
for w in cfg3 cfg2; do
  MSBWT_FINAL_DEEP=$deep timeout 600 python tools/pack_ab.py --workload $w > gpurun_out/r2i_${w}_deep$deep.json 2> gpurun_out/r2i_${w}.err
  echo "$w deep=$deep rc=$?"; cat gpurun_out/r2i_${w}_deep$deep.json
done
done
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2i_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2i_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2i_ncu.log
