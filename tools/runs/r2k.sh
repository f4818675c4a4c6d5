#!/bin/bash
# 2-GPU call: the in-library dispatcher on two devices (parity), the bench under torchrun at N = 2, cfg5 layout variants on GPU 0
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2k_smi.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi_device.py tests/test_gpu_parity.py::test_multi_device_split_matches_single -m gpu -q > gpurun_out/r2k_multi_device_pytest.log 2>&1
echo "multi rc=$?"; tail -8 gpurun_out/r2k_multi_device_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2k_bench_n2.json 2> gpurun_out/r2k_bench_n2.err
echo "bench n2 rc=$?"; grep -E "rank|e2e|parity|Error|error" gpurun_out/r2k_bench_n2.err | tail -25; head -c 700 gpurun_out/r2k_bench_n2.json; echo
for v in "13 0" "12 24"; do
  set -- $v
  MSBWT_FINAL_LINES_LOG2=$1 MSBWT_OCT_BUCKET_SHIFT=$2 timeout 600 python tools/pack_ab.py --workload cfg5 > gpurun_out/r2k_cfg5_lb$1_b$2.json 2> gpurun_out/r2k_cfg5.err
  echo "cfg5 lb=$1 b=$2 rc=$?"; cat gpurun_out/r2k_cfg5_lb$1_b$2.json
done
