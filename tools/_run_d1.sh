timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/d1_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/d1_pytest.log
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload ${WL:-cfg3} --steps ${ST:-5} > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "$tag rc=$?"; tail -2 gpurun_out/$tag.err; }
run d1_cfg3 A=1
WL=cfg2 ST=10 run d1_cfg2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/d1_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value %.4g ms %.3f kernel_ms %.3f e2e %.4g pile %.4g acc/q %.2f acc/s %.3g frac %.3f idx %.1f GB ts %s shift %s share %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms"], d["e2e"]["value"], d["e2e_pileup"]["value"], r["index_accesses_per_query"], r["index_accesses_per_s"], r["frac"], d["config"]["index_bytes"]/1e9, d["config"]["suffix_table_s"], r.get("oct_bucket_shift"), r.get("oct_overflow_position_share")))
    except Exception as e: print(f, "failed", e)
PY
