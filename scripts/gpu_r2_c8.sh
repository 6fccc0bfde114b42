#!/bin/bash
# Round 2, 8-GPU call: bench.py at N=8 with exchange / threshold / repair-round variants, then N=4.
mkdir -p gpurun_out
run() { # name, nproc, extra args...
  name=$1; NG=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu --no-probe "$@" > gpurun_out/c8_$name.json 2> gpurun_out/c8_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c8_$name.json") if l.startswith("{")][-1])
    print("$name N=%d qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f e2e_ids %.0f | pruned qps %.0f ms %.2f identical %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d["pruned"]["value"], d["pruned"]["ms_per_step"], d["pruned"]["results_identical"], d.get("sharded_breakdown_ms_per_call")))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c8_$name.err").read()[-2500:])
PY
}
run n8_default 8
run n8_nothr 8 --no-thr-exchange
run n8_allgather 8 --exchange allgather
BB25_REPAIR_ROUNDS=1 run n8_rr1 8
BB25_GROUPS=2 run n8_g2 8
run n4_default 4
