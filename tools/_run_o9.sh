timeout 900 python tools/ksweep.py > gpurun_out/o9_ksweep.json 2> gpurun_out/o9_ksweep.err; echo "ksweep rc=$?"; tail -3 gpurun_out/o9_ksweep.err
timeout 1200 python bench.py --workload cfg5 --steps 3 > gpurun_out/o9_cfg5.json 2> gpurun_out/o9_cfg5.err; echo "cfg5 rc=$?"; tail -6 gpurun_out/o9_cfg5.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/o9_ref.json 2> gpurun_out/o9_ref.err; echo "ref rc=$?"; head -c 700 gpurun_out/o9_ref.json
python - <<'PY'
import json
try:
    d=json.load(open("gpurun_out/o9_ksweep.json"))
    for r in d["results"]: print("k",r["k"],"ms",round(r["ms_ungrouped"],3),"q/s %.3g"%r["queries_per_s_ungrouped"])
except Exception as e: print("ksweep failed",e)
try:
    d=json.load(open("gpurun_out/o9_cfg5.json")); r=d["roofline"]
    print("cfg5 value %.4g ms %.3f kernel_ms %.3f e2e %.4g idx %.1f GB oct %s shift %s share %.4f" % (d["value"], d["ms_per_step"], r["kernel_ms"], d["e2e"]["value"], d["config"]["index_bytes"]/1e9, d["config"].get("oct_index"), r.get("oct_bucket_shift"), r.get("oct_overflow_position_share")))
except Exception as e: print("cfg5 failed",e)
PY
