#!/bin/bash
# where the end-to-end sharded step spends its wall time (N = number of visible GPUs)
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
BB25_BENCH_TRACE_E2E=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu --no-probe > gpurun_out/trace_n$NG.json 2> gpurun_out/trace_n$NG.err
python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/trace_n$NG.json") if l.startswith("{")][-1])
    print("N=%d qps %.0f ms/step %.2f e2e %.0f e2e_ids %.0f | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d.get("sharded_breakdown_ms_per_call")))
    for r,row in enumerate(d["e2e_trace"]["ranks"]): print(r, [round(x,2) for x in row])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/trace_n$NG.err").read()[-2500:])
PY
