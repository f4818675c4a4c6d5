#!/bin/bash
# two GPUs, final build: the in-library dispatcher tests and the bench at N = 2
mkdir -p gpurun_out
date +%T
timeout 600 python -m pytest tests/test_gpu_concurrency.py tests/test_gpu_multi_device.py "tests/test_gpu_parity.py::test_multi_device_split_matches_single" -m gpu -q > gpurun_out/r2z_pytest_2gpu.log 2>&1
echo "pytest rc=$?"; date +%T; tail -4 gpurun_out/r2z_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2z_bench_n2.json 2> gpurun_out/r2z_bench_n2.err
echo "bench n2 rc=$?"; date +%T; cut -c1-330 gpurun_out/r2z_bench_n2.json; python -c "
import json
d=json.loads(open('gpurun_out/r2z_bench_n2.json').read().strip().splitlines()[-1])
print('value %.2f G q/s  e2e %.2f G'%(d['value']/1e9, d['e2e']['value']/1e9), {k:(round(v['value']/1e9,2) if isinstance(v,dict) and 'value' in v else None) for k,v in d['e2e'].items() if isinstance(v,dict)})
"
