#!/bin/bash
mkdir -p gpurun_out
date +%T
MSBWT_TRACE=1 timeout 420 python -m pytest "tests/test_gpu_parity.py::test_wide_index_beyond_2_pow_32_symbols" -m gpu -q -x -s --durations=3 > gpurun_out/r2s_pytest_wide_big.log 2>&1
echo "big rc=$?"; date +%T; tail -15 gpurun_out/r2s_pytest_wide_big.log
timeout 400 python -m pytest tests/test_gpu_oct_index.py tests/test_gpu_final_step.py -m gpu -q --durations=5 > gpurun_out/r2s_pytest_wide_small.log 2>&1
echo "small rc=$?"; date +%T; tail -25 gpurun_out/r2s_pytest_wide_small.log
timeout 300 python tools/ksweep.py > gpurun_out/r2s_ksweep_cfg4.json 2> gpurun_out/r2s_ksweep.err
echo "ksweep rc=$?"; date +%T; python -c "
import json
d=json.load(open('gpurun_out/r2s_ksweep_cfg4.json'))
for r in d['results']: print(r['k'], round(r['ms'],3), round(r['ms_pack_stage'],3), round(r['queries_per_s']/1e9,2), round(r['line_fills_per_query'],2))
"
