#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2ad_bench.json 2> gpurun_out/r2ad_bench.err
echo "bench rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/r2ad_bench.json').read().strip().splitlines()[-1])
print(d['value']/1e9, d['e2e']['value']/1e9, d['clocks'], d['roofline']['frac'], d['gpu_launches'])
"
