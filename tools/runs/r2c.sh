#!/bin/bash
# round-2 third call: the one-request kernel (final_kernels.cu): parity suite, A/B against the general kernels, new bench.py
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -15 gpurun_out/r2c_pytest_gpu.log
for w in cfg3 cfg2; do
  for fast in 0 1; do
    MSBWT_FINAL_FAST=$fast timeout 600 python -X faulthandler tools/pack_ab.py --workload $w > gpurun_out/r2c_${w}_fast$fast.json 2> gpurun_out/r2c_${w}_fast$fast.err
    echo "$w fast=$fast rc=$?"; cat gpurun_out/r2c_${w}_fast$fast.json; tail -3 gpurun_out/r2c_${w}_fast$fast.err
  done
done
timeout 900 python bench.py --workload cfg2 --steps 3 > gpurun_out/r2c_bench_cfg2.json 2> gpurun_out/r2c_bench_cfg2.err
echo "bench cfg2 rc=$?"; tail -5 gpurun_out/r2c_bench_cfg2.err; head -c 1500 gpurun_out/r2c_bench_cfg2.json; echo
timeout 900 python bench.py --workload cfg3 --steps 5 > gpurun_out/r2c_bench_cfg3.json 2> gpurun_out/r2c_bench_cfg3.err
echo "bench cfg3 rc=$?"; tail -12 gpurun_out/r2c_bench_cfg3.err; head -c 1500 gpurun_out/r2c_bench_cfg3.json; echo
