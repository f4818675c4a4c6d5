#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv > gpurun_out/r2o_smi.txt 2>&1
nvidia-smi topo -m > gpurun_out/r2o_topo.txt 2>&1
( time timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2o_bench_n8.json 2> gpurun_out/r2o_bench_n8.err ) 2> gpurun_out/r2o_time.txt
echo "bench n8 rc=$?"; tail -3 gpurun_out/r2o_time.txt; grep -E "rank 0|e2e|Error|error|Traceback" gpurun_out/r2o_bench_n8.err | tail -30; head -c 500 gpurun_out/r2o_bench_n8.json; echo
