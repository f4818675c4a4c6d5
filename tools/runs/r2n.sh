#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_concurrency.py tests/test_gpu_multi_device.py -m gpu -q > gpurun_out/r2n_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r2n_pytest.log
( time MSBWT_BENCH_CFG5_ANY_N=1 timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2n_bench_n2.json 2> gpurun_out/r2n_bench_n2.err ) 2> gpurun_out/r2n_time.txt
echo "bench n2 rc=$?"; tail -3 gpurun_out/r2n_time.txt; grep -E "rank|e2e|parity|Error|error|Traceback" gpurun_out/r2n_bench_n2.err | tail -30; head -c 500 gpurun_out/r2n_bench_n2.json; echo
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2n_ref_n2.json 2> gpurun_out/r2n_ref_n2.err ) 2>> gpurun_out/r2n_time.txt
echo "ref n2 rc=$?"; head -c 300 gpurun_out/r2n_ref_n2.json; echo
