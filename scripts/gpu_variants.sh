#!/bin/bash
# A/B of tuning builds (scripts/build_variants.sh) on the full bench (no CPU leg).
# VARIANTS="name[:ENV=val,...]" ...; the first one also runs the GPU parity tests.
mkdir -p gpurun_out
first=1
for spec in $VARIANTS; do
  name=${spec%%:*}; envs=""
  [[ "$spec" == *:* ]] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  lib=$PWD/build_variants/libbb25_$name.so
  echo "== $name $envs"
  if [ $first = 1 ]; then
    env BB25_LIB=$lib $envs timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
    first=0
  fi
  tag=$(echo "$spec" | tr ':=,' '___')
  env BB25_LIB=$lib $envs timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_$tag.json").read().strip().splitlines()[-1])
    p=d.get("pruned") or {}
    print("exhaustive qps %.0f kernel_ms %.1f e2e %.0f reruns %.0f | pruned qps %.0f kernel_ms %.1f identical %s skipped %.0f ms-units %.0f" % (d["value"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["config"]["threshold_reruns_per_step"], p.get("value",0), p.get("kernel_ms_per_step",0), p.get("results_identical"), p.get("units_skipped_per_step",0), p.get("units_maxscore_per_step",0)))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/var_$tag.err").read()[-1500:])
PY
done
