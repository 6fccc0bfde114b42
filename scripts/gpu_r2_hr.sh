#!/bin/bash
# fp16 upper-bound rows on/off at shard-sized corpora (what N = 2 / 8 leave per GPU), exhaustive + pruned
mkdir -p gpurun_out
for docs in 4400000 1100000; do for h in 1 0; do
  BB25_HALF_ROWS=$h python bench.py --docs $docs --steps 5 --warmup 3 --no-cpu --no-probe > gpurun_out/hr_${docs}_$h.json 2> gpurun_out/hr_${docs}_$h.err
  python - $docs $h <<'PY'
import json,sys
d=json.loads([l for l in open("gpurun_out/hr_%s_%s.json"%(sys.argv[1],sys.argv[2])) if l.startswith("{")][-1]); p=d["pruned"]
print("docs %s half %s: exh qps %.0f ms %.2f kernel %.2f | pruned %.0f kernel %.2f" % (sys.argv[1], sys.argv[2], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], p["value"], p["kernel_ms_per_step"]))
PY
done; done
