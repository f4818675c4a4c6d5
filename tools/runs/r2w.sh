#!/bin/bash
mkdir -p gpurun_out
date +%T
timeout 300 python tools/ksweep.py > gpurun_out/r2w_ksweep_cfg4.json 2> gpurun_out/r2w_ksweep.err
echo "ksweep rc=$?"; date +%T; python -c "
import json
d=json.load(open('gpurun_out/r2w_ksweep_cfg4.json'))
for r in d['results']: print(r['k'], round(r['ms'],3), round(r['ms_pack_stage'],3), round(r['queries_per_s']/1e9,2), round(r['line_fills_per_query'],2))
"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:count_kmers_oct_kernel -s 3 -c 1 -o gpurun_out/r2w_oct_k63 -f python tools/pack_ab.py --workload cfg3 --iters 2 --watchdog 300 --k 63 --n 20000000 > gpurun_out/r2w_ncu.log 2>&1
echo "ncu rc=$?"; date +%T; tail -3 gpurun_out/r2w_ncu.log; ls -la gpurun_out/r2w_oct_k63.ncu-rep
