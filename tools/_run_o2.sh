timeout 1200 python -m pytest tests/test_gpu_oct_index.py -x -q > gpurun_out/o2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/o2_pytest.log
timeout 600 python bench.py --workload cfg3 --steps 5 > gpurun_out/o2_cfg3.json 2> gpurun_out/o2_cfg3.err; echo "cfg3 rc=$?"; tail -4 gpurun_out/o2_cfg3.err
MSBWT_LIBRARY_PATH=$PWD/build/variants/lib_oct3.so timeout 600 python bench.py --workload cfg3 --steps 5 > gpurun_out/o2_cfg3_ctas3.json 2> gpurun_out/o2_cfg3_ctas3.err; echo "cfg3 ctas3 rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 10 > gpurun_out/o2_cfg2.json 2> gpurun_out/o2_cfg2.err; echo "cfg2 rc=$?"
python - <<'PY'
import json
for f in ("o2_cfg3","o2_cfg3_ctas3","o2_cfg2"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); r=d["roofline"]
        print(f, "value %.4g ms %.3f kernel_ms %.3f e2e %.4g acc/q %.2f acc/s %.3g frac %.3f idx %.1f GB oct %s shift %s ovf %s share %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms"], d["e2e"]["value"], r["index_accesses_per_query"], r["index_accesses_per_s"], r["frac"], d["config"]["index_bytes"]/1e9, d["config"].get("oct_index"), r.get("oct_bucket_shift"), r.get("oct_overflow_lines"), r.get("oct_overflow_position_share")))
    except Exception as e: print(f, "failed", e)
PY
