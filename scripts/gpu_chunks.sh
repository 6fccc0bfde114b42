#!/bin/bash
# N-GPU run of bench.py for several --shard-chunks settings
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for c in ${CHUNKS:-1 2 4}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29600+c)) bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu --shard-chunks $c > gpurun_out/chunks_$c.json 2> gpurun_out/chunks_$c.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/chunks_$c.json") if l.startswith("{")][-1])
    print("chunks=$c N=%d qps %.0f ms/step %.1f kernel_ms %.1f e2e %.0f | pruned qps %.0f e2e %.0f" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["pruned"]["value"], d["pruned"]["e2e_value"]))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/chunks_$c.err").read()[-1500:])
PY
done
