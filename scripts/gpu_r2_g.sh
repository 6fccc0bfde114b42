#!/bin/bash
# two-kernel fused traversal: parity tests, then config 5 at 6 M docs per tuning build, then full size
mkdir -p gpurun_out
echo "== fused tests"; timeout 1200 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout=1000 --deselect tests/test_gpu_fused.py::test_config5_full_size_50m_docs > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/g_tests.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1]); p=d["pruned"]; n=d.get("pruned_without_essential_evaluation") or {}
    print("%s: pruned qps %.0f ms %.1f kernel %.1f sparse %.0f skipped %.0f ident %s | nosparse %.0f | exh %.0f | parity %s" % (sys.argv[1], p["value"], p["ms_per_step"], p["kernel_ms_per_step"], p["units_by_essential_postings_per_step"], p["units_skipped_per_step"], p["results_identical_to_exhaustive"], n.get("value",0), d["exhaustive"]["value"], d.get("parity_spot_check")))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
for v in def $VARIANTS; do
  lib=$PWD/build_variants/libbb25_$v.so; [ $v = def ] && lib=$PWD/bayesian_bm25_b200/libbb25.so
  BB25_LIB=$lib timeout 600 python bench.py --config 5 --docs 6000000 --queries 2000 --steps 3 --warmup 1 --no-cpu > gpurun_out/g_c5s_$v.json 2> gpurun_out/g_c5s_$v.err; show gpurun_out/g_c5s_$v.json
done
if [ -z "$SKIP_FULL" ]; then
timeout 1200 python bench.py --config 5 --steps 2 --warmup 1 > gpurun_out/g_c5.json 2> gpurun_out/g_c5.err; echo "c5 rc=$?"; tail -2 gpurun_out/g_c5.err; show gpurun_out/g_c5.json
fi
