#!/bin/bash
# A/B of the split evaluation (register sums for documents matching frequent terms only).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for s in ${SPLIT_CONFIGS:-1 0}; do
  echo "== split $s"
  BB25_SPLIT=$s timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/split_$s.json 2> gpurun_out/split_$s.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/split_$s.json").read().strip().splitlines()[-1])
    print("qps %.0f ms/step %.1f kernel_ms %.1f e2e %.0f reruns %.0f | pruned %s | spot %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["config"]["threshold_reruns_per_step"], d.get("pruned"), d.get("parity_spot_check")))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/split_$s.err").read()[-1500:])
PY
done
