#!/bin/bash
# Round 2, call C (2+ GPUs): NCCL sharded parity test, then bench.py at N = visible GPUs with exchange variants.
mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l)
echo "== nccl test ($NG GPUs)"; timeout 900 python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q -x -s --timeout=800 > gpurun_out/c_nccl_test.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/c_nccl_test.log
run() { # name, extra args...
  name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $NG --steps 5 --warmup 3 --no-cpu --no-probe "$@" > gpurun_out/c_$name.json 2> gpurun_out/c_$name.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/c_$name.json") if l.startswith("{")][-1])
    print("$name N=%d qps %.0f ms/step %.2f kernel_ms %.2f e2e %.0f e2e_ids %.0f | pruned qps %.0f identical %s | %s | %s" % (d["n_gpus"], d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_per_step"], d["e2e"]["value"], d["e2e"]["retrieve_ids_value"], d["pruned"]["value"], d["pruned"]["results_identical"], d.get("sharded_breakdown_ms_per_call"), d["config"]["parallelism"]))
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/c_$name.err").read()[-2500:])
PY
}
run sliced_thr
run sliced_nothr --no-thr-exchange
run allgather_thr --exchange allgather
run allgather_nothr --exchange allgather --no-thr-exchange
BB25_SYMM=0 run slicednccl_thr
