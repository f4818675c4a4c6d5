#!/bin/bash
# the WIDE kernels of the FINAL build on the configs[2] index cut into superblocks
mkdir -p gpurun_out
timeout 200 python tools/pack_ab.py --workload cfg3 --iters 10 --watchdog 120 --superblock-shift 20 --also 43:100000000,63:10000000,101:10000000 > gpurun_out/r2al_wide_cfg3.jsonl 2> gpurun_out/r2al_wide.err
echo "wide rc=$?"; python -c "
import json
for l in open('gpurun_out/r2al_wide_cfg3.jsonl'):
    d=json.loads(l); print('   wide k %d n %d: search %.3f ms pack %.3f ms  %.2f G q/s present %d checksum %d'%(d['k'],d['queries'],d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['checksum']))
"
