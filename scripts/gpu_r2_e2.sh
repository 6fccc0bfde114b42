#!/bin/bash
# config 5 line with the essential-posting evaluation A/B
mkdir -p gpurun_out
echo "== config 5"; timeout 1200 python bench.py --config 5 --steps ${STEPS:-2} --warmup 1 > gpurun_out/e_c5.json 2> gpurun_out/e_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/e_c5.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/e_c5.json") if l.startswith("{")][-1])
    print("qps %.0f ms %.1f" % (d["value"], d["ms_per_step"]))
    print("pruned", d["pruned"])
    print("nosparse", d.get("pruned_without_essential_evaluation"))
    print("exh", d["exhaustive"]); print("parity", d.get("parity_spot_check"), d.get("parity_spot_check_detail"))
except Exception as e:
    print("FAILED", e)
PY
