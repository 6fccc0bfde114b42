#!/bin/bash
# Round 2, call A: parity suite on the sync-free pipeline, L2/HBM read peaks, bench line with the live
# ncu traffic probe, repair-round A/B.  Everything under `timeout`; logs in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== peaks"; timeout 300 python scripts/peaks.py > gpurun_out/peaks_l2.json 2> gpurun_out/peaks.err; echo "rc=$?"; cat gpurun_out/peaks_l2.json | head -30
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout=1200 -x ${PYTEST_ARGS} > gpurun_out/a_tests.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/a_tests.log
echo "== bench"; timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/a_bench.json; tail -5 gpurun_out/a_bench.err
for r in 0 1; do
  echo "== repair rounds $r"
  BB25_REPAIR_ROUNDS=$r timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --no-probe > gpurun_out/a_bench_rr$r.json 2> gpurun_out/a_bench_rr$r.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/a_bench_rr$r.json") if l.startswith("{")][-1])
    print("rr=$r qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f e2e_ids %.0f syncs %.1f repaired %.1f dense %.1f | pruned qps %.0f identical %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d["config"]["host_syncs_per_step"], d["config"]["repaired_queries_per_step"], d["config"]["dense_fallback_queries_per_step"], d["pruned"]["value"], d["pruned"]["results_identical"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/a_bench_rr$r.err").read()[-1500:])
PY
done
