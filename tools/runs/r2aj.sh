#!/bin/bash
# FINAL build of round 2: full GPU suite, smoke, default bench, k-sweep
mkdir -p gpurun_out
date +%T
timeout 1100 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/r2aj_pytest_gpu.log 2>&1
echo "pytest rc=$?"; date +%T; tail -10 gpurun_out/r2aj_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2aj_smoke.log 2>&1
echo "smoke rc=$?"; tail -1 gpurun_out/r2aj_smoke.log
timeout 600 python bench.py > gpurun_out/r2aj_bench.json 2> gpurun_out/r2aj_bench.err
echo "bench rc=$?"; date +%T; python -c "
import json
d=json.loads(open('gpurun_out/r2aj_bench.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['e2e']['value']/1e9, d['clocks'], d['roofline']['frac'], d['gpu_launches'])
"
timeout 300 python tools/ksweep.py > gpurun_out/r2aj_ksweep_cfg4.json 2> gpurun_out/r2aj_ksweep.err
echo "ksweep rc=$?"; date +%T; python -c "
import json
d=json.load(open('gpurun_out/r2aj_ksweep_cfg4.json'))
for r in d['results']: print(r['k'], round(r['ms'],3), round(r['ms_pack_stage'],3), round(r['queries_per_s']/1e9,2), round(r['line_fills_per_query'],2))
"
