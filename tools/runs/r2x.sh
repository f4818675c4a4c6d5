#!/bin/bash
mkdir -p gpurun_out
date +%T
timeout 600 python -m pytest tests/test_gpu_final_step.py tests/test_gpu_oct_index.py tests/test_gpu_final_fast.py tests/test_gpu_fused.py tests/test_gpu_full_size.py "tests/test_gpu_parity.py::test_wide_index_beyond_2_pow_32_symbols" -m gpu -q -x > gpurun_out/r2x_pytest.log 2>&1
echo "pytest rc=$?"; date +%T; tail -4 gpurun_out/r2x_pytest.log
timeout 200 python tools/pack_ab.py --workload cfg3 --iters 10 --watchdog 60 --k 63 --n 10000000 --also 63:70000000,101:10000000,43:100000000,53:40000000 > gpurun_out/r2x_long.jsonl 2> gpurun_out/r2x_long.err
echo "pack_ab rc=$?"; date +%T; python -c "
import json
for l in open('gpurun_out/r2x_long.jsonl'):
    d=json.loads(l); print('   k %d n %d: search %.3f ms pack %.3f ms  %.2f G q/s present %d checksum %d'%(d['k'],d['queries'],d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['checksum']))
"
