#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err ) 2> gpurun_out/r2l_bench.time
echo "bench rc=$?"; cat gpurun_out/r2l_bench.time | tail -3; grep -E "rank 0|e2e|cpu\]" gpurun_out/r2l_bench.err | tail -30; head -c 400 gpurun_out/r2l_bench.json; echo
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2l_bench_ref.json 2> gpurun_out/r2l_bench_ref.err ) 2> gpurun_out/r2l_ref.time
echo "ref rc=$?"; tail -3 gpurun_out/r2l_ref.time; tail -3 gpurun_out/r2l_bench_ref.err; head -c 900 gpurun_out/r2l_bench_ref.json; echo
timeout 300 python tools/gather_bulk.py > gpurun_out/r2l_gather_bulk.json 2> gpurun_out/r2l_gather_bulk.err; echo "bulk rc=$?"; cat gpurun_out/r2l_gather_bulk.json
timeout 900 python tools/ksweep.py > gpurun_out/r2l_ksweep_cfg4.json 2> gpurun_out/r2l_ksweep.err; echo "ksweep rc=$?"; tail -5 gpurun_out/r2l_ksweep.err | cut -c1-600
