#!/bin/bash
# A/B of traversal-kernel variants / tile sizes on the full bench (no CPU leg).
mkdir -p gpurun_out
CFGS=${AB_CONFIGS:-0:16384 1:16384 2:16384 3:16384 3:32768 3:8192}
for cfg in $CFGS; do
  v=${cfg%%:*}; t=${cfg##*:}
  echo "== variant $v tile $t"
  BB25_VARIANT=$v BB25_TILE_DOCS=$t timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ab_${v}_${t}.json 2> gpurun_out/ab_${v}_${t}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${v}_${t}.json").read().strip().splitlines()[-1])
    print("qps %.0f ms/step %.1f kernel_ms %.1f frac %.3f e2e %.0f launches %d reruns %.0f" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["config"]["threshold_reruns_per_step"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/ab_${v}_${t}.err").read()[-1500:])
PY
done
