#!/bin/bash
# full GPU test suite + N=1 bench line (+ optional variants through BB25_LIB)
mkdir -p gpurun_out
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q --timeout=1500 ${PYTEST_ARGS} > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/d_tests.log
echo "== bench"; timeout 1500 python bench.py --steps ${STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d_bench.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open("gpurun_out/d_bench.json") if l.startswith("{")][-1])
    r=d["roofline"]
    print("qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f e2e_ids %.0f | pruned qps %.0f kernel %.2f e2e %.0f identical %s | bound %s frac %s fractions %s | parity %s" % (d["value"], d["ms_per_step"], r["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d["pruned"]["value"], d["pruned"]["kernel_ms_per_step"], d["pruned"]["e2e_value"], d["pruned"]["results_identical"], r.get("bound"), r.get("frac"), r.get("fractions"), d.get("parity_spot_check")))
except Exception as e:
    print("FAILED", e)
PY
