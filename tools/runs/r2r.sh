#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_oct_index.py tests/test_gpu_final_fast.py tests/test_gpu_final_step.py tests/test_gpu_fused.py -m gpu -q -x > gpurun_out/r2r_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2r_pytest.log
timeout 600 python tools/pack_ab.py --workload cfg3 > gpurun_out/r2r_cfg3.json 2> gpurun_out/r2r.err; cat gpurun_out/r2r_cfg3.json
timeout 600 python tools/pack_ab.py --workload cfg3 --n 12500000 > gpurun_out/r2r_cfg3_n12m.json 2>> gpurun_out/r2r.err; cat gpurun_out/r2r_cfg3_n12m.json
timeout 600 python tools/pack_ab.py --workload cfg3 --k 63 --n 10000000 > gpurun_out/r2r_cfg3_k63.json 2>> gpurun_out/r2r.err; cat gpurun_out/r2r_cfg3_k63.json
