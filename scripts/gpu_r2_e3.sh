#!/bin/bash
# fused parity (without the 50 M test) + config 5 at 6 M docs and at full size
mkdir -p gpurun_out
echo "== fused tests"; timeout 1200 python -m pytest tests/test_gpu_fused.py -m gpu -q -x --timeout=1000 --deselect tests/test_gpu_fused.py::test_config5_full_size_50m_docs > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/e_tests.log
show() { python - "$1" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print("qps %.0f ms %.1f" % (d["value"], d["ms_per_step"]))
    print("pruned", d["pruned"])
    print("nosparse", d.get("pruned_without_essential_evaluation"))
    print("exh", d["exhaustive"]); print("parity", d.get("parity_spot_check"), d.get("parity_spot_check_detail"))
except Exception as e:
    print("FAILED", e)
PY
}
echo "== config 5 @6M"; timeout 600 python bench.py --config 5 --docs 6000000 --queries 2000 --steps 2 --warmup 1 --no-cpu > gpurun_out/e_c5s.json 2> gpurun_out/e_c5s.err; echo "rc=$?"; tail -3 gpurun_out/e_c5s.err; show gpurun_out/e_c5s.json
if [ -z "$SKIP_FULL" ]; then
echo "== config 5"; timeout 1200 python bench.py --config 5 --steps 2 --warmup 1 > gpurun_out/e_c5.json 2> gpurun_out/e_c5.err; echo "c5 rc=$?"; tail -3 gpurun_out/e_c5.err; show gpurun_out/e_c5.json
fi
