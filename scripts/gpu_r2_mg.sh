#!/bin/bash
# multi-GPU evidence: NCCL sharded parity test (2+ GPUs), then the bench line at N = visible GPUs
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
if [ -n "$RUN_TEST" ]; then
echo "== nccl test ($NG GPUs)"; timeout 900 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q -x -s --timeout=800 > gpurun_out/mg_nccl_test_n$NG.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/mg_nccl_test_n$NG.log
fi
BB25_BENCH_TRACE_E2E=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29713 bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu > gpurun_out/mg_n$NG.json 2> gpurun_out/mg_n$NG.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/mg_n$NG.json") if l.startswith("{")][-1])
    print("N=%d qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f e2e_ids %.0f | pruned qps %.0f e2e %.0f identical %s | %s | parity %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d["pruned"]["value"], d["pruned"]["e2e_value"], d["pruned"]["results_identical"], d.get("sharded_breakdown_ms_per_call"), d.get("parity_spot_check")))
    for r,row in enumerate((d.get("e2e_trace") or {}).get("ranks", [])[:2]): print(r, [round(x,2) for x in row])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/mg_n$NG.err").read()[-2500:])
PY
