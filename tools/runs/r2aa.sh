#!/bin/bash
# the WIDE one-request kernel: parity tests and the configs[2] index cut into superblocks
mkdir -p gpurun_out
date +%T
timeout 600 python -m pytest tests/test_gpu_final_fast.py tests/test_gpu_oct_index.py tests/test_gpu_final_step.py tests/test_gpu_u64_kmers.py tests/test_gpu_host_pack.py "tests/test_gpu_parity.py::test_wide_index_beyond_2_pow_32_symbols" -m gpu -q -x > gpurun_out/r2aa_pytest.log 2>&1
echo "pytest rc=$?"; date +%T; tail -4 gpurun_out/r2aa_pytest.log
timeout 300 python tools/pack_ab.py --workload cfg3 --iters 10 --watchdog 120 --superblock-shift 20 --also 32:100000000,63:10000000 > gpurun_out/r2aa_wide_cfg3.jsonl 2> gpurun_out/r2aa_wide.err
echo "wide rc=$?"; date +%T; python -c "
import json
for l in open('gpurun_out/r2aa_wide_cfg3.jsonl'):
    d=json.loads(l); print('   wide k %d n %d: search %.3f ms pack %.3f ms  %.2f G q/s present %d checksum %d index %.1f GB'%(d['k'],d['queries'],d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['checksum'],d['index_bytes']/1e9))
"
