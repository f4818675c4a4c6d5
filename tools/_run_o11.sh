timeout 1500 python -m pytest tests/test_gpu_callers.py -x -q > gpurun_out/o11_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/o11_pytest.log
timeout 900 python bench.py --workload cfg2 --steps 10 > gpurun_out/o11_cfg2.json 2> gpurun_out/o11_cfg2.err; echo "cfg2 rc=$?"; tail -3 gpurun_out/o11_cfg2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/o11_cfg2.json")); print("value %.4g e2e %.4g pileup %s" % (d["value"], d["e2e"]["value"], d["e2e_pileup"]))
PY
