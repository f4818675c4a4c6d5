#!/bin/bash
# ncu launch list of the default bench command on the final build (our kernels only)
mkdir -p gpurun_out
OURS='regex:pack_seed|seed_packed|seed_u64|count_kmers|constrain_ranges|narrow_counts|expand_read|sum_strands|gather'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv --log-file gpurun_out/r2ac_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2ac_ncu_bench.log 2>&1
echo "ncu launch list rc=$?"; date +%T; grep -c "pack_seed_final_kernel" gpurun_out/r2ac_launches.csv; wc -l gpurun_out/r2ac_launches.csv
