#!/bin/bash
# line requests exchanged through the row padding instead of shuffles: stress on the shapes that exposed the convergence bug, parity, timings
mkdir -p gpurun_out
date +%T
run() { # k n iters
timeout 120 python tools/pack_ab.py --workload cfg3 --iters $3 --watchdog 110 --postmortem 8 --prefill 0 --k $1 --n $2 > gpurun_out/r2ai_k$1_n$2.jsonl 2> gpurun_out/r2ai_k$1_n$2.err
echo "k=$1 n=$2 iters=$3 rc=$?"; python -c "
import json,sys
for l in open('gpurun_out/r2ai_k$1_n$2.jsonl'):
    d=json.loads(l); print('   search %.3f ms pack %.3f ms  %.2f G q/s present %d of %d checksum %d'%(d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['queries'],d['checksum']))
"; grep "postmortem\|illegal" gpurun_out/r2ai_k$1_n$2.err | head -3 | cut -c1-300
}
run 43 100000000 40
run 63 70000000 30
run 101 40000000 12
run 63 10000000 12
date +%T
timeout 700 python -m pytest tests/test_gpu_oct_index.py tests/test_gpu_final_step.py tests/test_gpu_final_fast.py tests/test_gpu_fused.py tests/test_gpu_full_size.py tests/test_gpu_quad_index.py "tests/test_gpu_parity.py::test_wide_index_beyond_2_pow_32_symbols" -m gpu -q -x > gpurun_out/r2ai_pytest.log 2>&1
echo "pytest rc=$?"; date +%T; tail -4 gpurun_out/r2ai_pytest.log
