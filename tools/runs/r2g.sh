#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_final_fast.py tests/test_gpu_final_step.py tests/test_gpu_u64_kmers.py tests/test_gpu_host_pack.py -m gpu -q -x > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2g_pytest.log
for w in cfg3 cfg2; do
  timeout 600 python tools/pack_ab.py --workload $w > gpurun_out/r2g_${w}.json 2> gpurun_out/r2g_${w}.err
  echo "$w rc=$?"; cat gpurun_out/r2g_${w}.json
done
timeout 300 python tools/pcie_probe.py > gpurun_out/r2g_pcie.json 2>&1; cat gpurun_out/r2g_pcie.json
timeout 600 python tools/group_bench.py --workload cfg3 > gpurun_out/r2g_group_cfg3.json 2> gpurun_out/r2g_group.err; echo "group rc=$?"; cat gpurun_out/r2g_group_cfg3.json; tail -3 gpurun_out/r2g_group.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2g_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2g_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2g_ncu.log
