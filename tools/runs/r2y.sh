#!/bin/bash
# FINAL single-GPU evidence set of round 2
mkdir -p gpurun_out
date +%T
timeout 1100 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/r2y_pytest_gpu.log 2>&1
echo "pytest rc=$?"; date +%T; tail -14 gpurun_out/r2y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2y_smoke.log
timeout 600 python bench.py > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err
echo "bench rc=$?"; date +%T; cut -c1-330 gpurun_out/r2y_bench.json
timeout 300 python tools/ksweep.py > gpurun_out/r2y_ksweep_cfg4.json 2> gpurun_out/r2y_ksweep.err
echo "ksweep rc=$?"; date +%T; python -c "
import json
d=json.load(open('gpurun_out/r2y_ksweep_cfg4.json'))
for r in d['results']: print(r['k'], round(r['ms'],3), round(r['ms_pack_stage'],3), round(r['queries_per_s']/1e9,2), round(r['line_fills_per_query'],2))
"
timeout 300 python tools/pack_ab.py --workload cfg3 --iters 10 --watchdog 120 --superblock-shift 20 --also 63:10000000,43:100000000 > gpurun_out/r2y_wide_cfg3.jsonl 2> gpurun_out/r2y_wide.err
echo "wide rc=$?"; date +%T; python -c "
import json
for l in open('gpurun_out/r2y_wide_cfg3.jsonl'):
    d=json.loads(l); print('   wide k %d n %d: search %.3f ms pack %.3f ms  %.2f G q/s present %d checksum %d index %.1f GB'%(d['k'],d['queries'],d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['checksum'],d['index_bytes']/1e9))
"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2y_bench_ref.json 2> gpurun_out/r2y_bench_ref.err
echo "ref rc=$?"; date +%T; cut -c1-300 gpurun_out/r2y_bench_ref.json
