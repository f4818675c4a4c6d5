#!/bin/bash
# ncu --set full of the FINAL oct kernel on 20 M 63-mers (after the two-phase scan and the request exchange)
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:count_kmers_oct_kernel -s 3 -c 1 -o gpurun_out/r2ah_oct_k63 -f python tools/pack_ab.py --workload cfg3 --iters 2 --watchdog 300 --k 63 --n 20000000 > gpurun_out/r2ah_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2ah_oct_k63.ncu-rep
