timeout 900 python -m pytest tests/test_gpu_fused.py -x -q > gpurun_out/f1_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/f1_pytest.log
run() { tag=$1; shift; env "$@" timeout 900 python bench.py --workload ${WL:-cfg3} --steps ${ST:-5} > gpurun_out/$tag.json 2> gpurun_out/$tag.err; echo "$tag rc=$?"; tail -2 gpurun_out/$tag.err; }
run f1_cfg3 A=1
WL=cfg2 ST=10 run f1_cfg2 A=1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/f1_*.json")):
    try:
        d=json.load(open(f)); r=d["roofline"]
        print(f, "value %.4g ms %.3f kernel_ms %.3f e2e %.4g pile %.4g acc/q %.2f acc/s %.3g frac %.3f fused %s" % (d["value"], d["ms_per_step"], r["kernel_ms"], d["e2e"]["value"], d["e2e_pileup"]["value"], r["index_accesses_per_query"], r["index_accesses_per_s"], r["frac"], r.get("fused")))
    except Exception as e: print(f, "failed", e)
PY
