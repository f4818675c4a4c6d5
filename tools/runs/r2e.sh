#!/bin/bash
# round-2 call e: header-hopping line scan in the one-request kernel: parity subset, timing, ncu capture, PCIe probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_final_fast.py tests/test_gpu_final_step.py tests/test_gpu_oct_index.py -m gpu -q -x > gpurun_out/r2e_pytest.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2e_pytest.log
for w in cfg3 cfg2; do
  timeout 600 python tools/pack_ab.py --workload $w > gpurun_out/r2e_${w}.json 2> gpurun_out/r2e_${w}.err
  echo "$w rc=$?"; cat gpurun_out/r2e_${w}.json
done
timeout 300 python tools/pcie_probe.py > gpurun_out/r2e_pcie.json 2>&1; cat gpurun_out/r2e_pcie.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2e_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2e_ncu.log
