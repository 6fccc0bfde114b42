#!/bin/bash
# A/B of tuning builds on the headline bench (no CPU leg, no probe): VARIANTS="name ..."
mkdir -p gpurun_out
for name in $VARIANTS; do
  lib=$PWD/build_variants/libbb25_$name.so
  env BB25_LIB=$lib timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-probe > gpurun_out/v_$name.json 2> gpurun_out/v_$name.err
  python - "$name" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/v_%s.json"%sys.argv[1]) if l.startswith("{")][-1]); p=d["pruned"]
    print("%-10s exh qps %.0f ms %.2f kernel %.2f e2e %.0f | pruned qps %.0f kernel %.2f ident %s" % (sys.argv[1], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], p["value"], p["kernel_ms_per_step"], p["results_identical"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/v_%s.err"%sys.argv[1]).read()[-800:])
PY
done
if [ -n "$NEHIST" ]; then
  env BB25_LIB=$PWD/build_variants/libbb25_nehist.so BB25_BENCH_AB_SPARSE=0 timeout 900 python bench.py --config 5 --steps 1 --warmup 1 --no-cpu > gpurun_out/v_nehist.json 2> gpurun_out/v_nehist.err
  python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/v_nehist.json") if l.startswith("{")][-1]); print("nehist", d["pruned"])
except Exception as e:
    print("nehist FAILED", e); print(open("gpurun_out/v_nehist.err").read()[-800:])
PY
fi
