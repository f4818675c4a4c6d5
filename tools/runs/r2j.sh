#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_final_fast.py -m gpu -q -x > gpurun_out/r2j_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2j_pytest.log
for deep in 1 0; do
  MSBWT_FINAL_DEEP=$deep timeout 600 python tools/pack_ab.py --workload cfg3 > gpurun_out/r2j_cfg3_deep$deep.json 2> gpurun_out/r2j_cfg3.err
  echo "cfg3 deep=$deep rc=$?"; cat gpurun_out/r2j_cfg3_deep$deep.json
done
timeout 600 python tools/pack_ab.py --workload cfg2 > gpurun_out/r2j_cfg2.json 2> gpurun_out/r2j_cfg2.err; cat gpurun_out/r2j_cfg2.json
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2j_final_cfg3 \
   python tools/pack_ab.py --workload cfg3 --iters 2 > gpurun_out/r2j_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r2j_ncu.log
MSBWT_TRACE=1 timeout 1200 python bench.py --workload cfg5 --steps 5 > gpurun_out/r2j_bench_cfg5.json 2> gpurun_out/r2j_bench_cfg5.err
echo "bench cfg5 rc=$?"; grep -E "msbwt|rank 0|e2e|cpu" gpurun_out/r2j_bench_cfg5.err | tail -30; head -c 600 gpurun_out/r2j_bench_cfg5.json; echo
