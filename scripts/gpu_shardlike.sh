#!/bin/bash
# Fixed per-step overheads at a shard-sized corpus (what N=8 leaves per GPU): bench lines for
# block-group counts, plus one launch list.
mkdir -p gpurun_out
DOCS=${SHARD_DOCS:-1100000}
for g in ${GROUPS_LIST:-3 2 1}; do
  BB25_GROUPS=$g python bench.py --docs $DOCS --steps 5 --warmup 3 --no-cpu > gpurun_out/shard_g$g.json 2> gpurun_out/shard_g$g.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/shard_g$g.json").read().strip().splitlines()[-1]); p=d["pruned"]
print("groups $g: exhaustive qps %.0f ms/step %.2f kernel %.2f reruns %.0f | pruned qps %.0f ms/step %.2f kernel %.2f identical %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["config"]["threshold_reruns_per_step"], p["value"], p["ms_per_step"], p["kernel_ms_per_step"], p["results_identical"]))
PY
done
true
true
