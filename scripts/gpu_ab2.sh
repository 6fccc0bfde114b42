#!/bin/bash
# A/B of traversal kernel families on the full bench (no CPU leg).
mkdir -p gpurun_out
CFGS=${AB2_CONFIGS:-block:0:6 block:1:6 block:2:5}
for cfg in $CFGS; do
  IFS=: read kern prune ctas <<< "$cfg"
  echo "== kernel $kern prune $prune ctas/SM $ctas"
  BB25_KERNEL=$kern BB25_PRUNE=$prune BB25_CTAS_PER_SM=$ctas timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/ab2_${kern}_${prune}_${ctas}.json 2> gpurun_out/ab2_${kern}_${prune}_${ctas}.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab2_${kern}_${prune}_${ctas}.json").read().strip().splitlines()[-1])
    print("qps %.0f ms/step %.1f kernel_ms %.1f frac %.3f e2e %.0f launches %d reruns %.0f | %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"], d["gpu_launches"], d["config"]["threshold_reruns_per_step"], d["config"].get("pruning")))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/ab2_${kern}_${prune}_${ctas}.err").read()[-1500:])
PY
done
