#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout=1000 > gpurun_out/h2_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/h2_tests.log
run() { # tag, env..., -- bench args
  tag=$1; shift
  env "$@" python bench.py --steps 4 --warmup 2 --no-cpu --no-probe $BARGS > gpurun_out/h_$tag.json 2> gpurun_out/h_$tag.err
  python - "$tag" <<'PY'
import json,sys
try:
    d=json.loads([l for l in open("gpurun_out/h_%s.json"%sys.argv[1]) if l.startswith("{")][-1]); p=d["pruned"]
    print("%-22s exh %.0f | pruned qps %.0f ms %.2f kernel %.2f e2e %.0f ident %s units %.0f skipped %.0f ms-units %.0f sparse %.0f routed %.0f" % (sys.argv[1], d["value"], p["value"], p["ms_per_step"], p["kernel_ms_per_step"], p["e2e_value"], p["results_identical"], p["units_per_step"], p["units_skipped_per_step"], p["units_maxscore_per_step"], p.get("units_by_essential_postings_per_step",-1), p["queries_routed_to_candidate_path_per_step"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/h_%s.err"%sys.argv[1]).read()[-600:])
PY
}
BARGS="--prune-level 3" run l3_sparse X=1
BARGS="--prune-level 2" run l2_sparse X=1
BARGS="--prune-level 3" run l3_sparse_rd128 BB25_ROUTE_DIV=128
BARGS="--prune-level 3" run l3_sparse_rd1024 BB25_ROUTE_DIV=1024
