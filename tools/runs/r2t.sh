#!/bin/bash
mkdir -p gpurun_out
date +%T
run() { # k n iters
timeout 120 python tools/pack_ab.py --workload cfg3 --iters $3 --watchdog 110 --postmortem 8 --prefill 0 --k $1 --n $2 > gpurun_out/r2u_k$1_n$2.jsonl 2> gpurun_out/r2u_k$1_n$2.err
echo "k=$1 n=$2 iters=$3 rc=$?"; date +%T; python -c "
import json,sys
for l in open('gpurun_out/r2u_k$1_n$2.jsonl'):
    d=json.loads(l); print('   search %.3f ms pack %.3f ms  %.2f G q/s present %d of %d checksum %d'%(d['search_ms_median'],d['pack_ms_median'],d['queries_per_s']/1e9,d['present'],d['queries'],d['checksum']))
"; grep "postmortem\|illegal" gpurun_out/r2u_k$1_n$2.err | head -4 | cut -c1-330
}
run 43 100000000 40
run 63 70000000 30
run 53 85000000 30
run 101 40000000 15
run 63 10000000 15
