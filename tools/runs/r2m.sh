#!/bin/bash
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py --smoke > gpurun_out/r2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2m_smoke.log
timeout 1700 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2m_pytest_gpu.log
timeout 400 python tools/ksweep.py > gpurun_out/r2m_ksweep_cfg4.json 2> gpurun_out/r2m_ksweep.err; echo "ksweep rc=$?"; tail -4 gpurun_out/r2m_ksweep.err | cut -c1-500
OURS='regex:pack_seed|seed_packed|seed_u64|count_kmers|constrain_ranges|narrow_counts|expand_read|sum_strands|gather'
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$OURS" -c 600 --csv \
    --log-file gpurun_out/r2m_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r2m_launches.log 2>&1
echo "launches rc=$?"; wc -l gpurun_out/r2m_launches.csv
for w in cfg5 cfg2; do
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:pack_seed_final" -s 3 -c 1 -f -o gpurun_out/r2m_final_$w \
   python tools/pack_ab.py --workload $w --iters 2 > gpurun_out/r2m_ncu_$w.log 2>&1
echo "ncu $w rc=$?"
done
