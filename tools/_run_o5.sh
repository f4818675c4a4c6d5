timeout 1200 python -m pytest tests/test_gpu_oct_index.py -x -q > gpurun_out/o5_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/o5_pytest.log
MSBWT_OCT_BUCKET_SHIFT=20 timeout 1200 ncu --set full --clock-control none --import-source on -k regex:count_kmers_oct -s 3 -c 1 \
        -f -o gpurun_out/o5_oct_cfg3 python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/o5_full_cfg3.log 2>&1
echo "full_cfg3 rc=$?"
