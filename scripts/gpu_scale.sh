#!/bin/bash
# Strong-scaling run on one box: bench.py at N = 1, 2, 4, 8 (as many GPUs as visible).
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
for n in ${SCALE_NS:-1 2 4 8}; do
  [ $n -gt $NG ] && break
  echo "== N=$n"
  if [ $n -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/scale_n$n.json") if l.startswith("{")][-1])
    print("N=%d qps %.0f ms/step %.1f kernel_ms %.1f e2e %.0f | pruned qps %.0f e2e %.0f identical %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["pruned"]["value"], d["pruned"]["e2e_value"], d["pruned"]["results_identical"], d.get("sharded_breakdown_ms_per_call")))
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/scale_n$n.err").read()[-2000:])
PY
done
