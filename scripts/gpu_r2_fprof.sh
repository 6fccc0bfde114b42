#!/bin/bash
# ncu --set full of the largest fused_block_kernel launch (group 3, pruning level 3) on a 6 M-document config-5 corpus
mkdir -p gpurun_out
A="--config 5 --docs 6000000 --queries 2000 --steps 1 --warmup 1 --no-cpu"
python bench.py $A > gpurun_out/fprof_plain.json 2> gpurun_out/fprof_plain.err; echo "plain rc=$?"
ncu --set full --clock-control none --import-source on -k regex:fused_block_kernel -s 8 -c 1 -f -o gpurun_out/prof_r02_fused_${TAG:-sparse} python bench.py $A > gpurun_out/fprof_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/fprof_ncu.log
ncu -i gpurun_out/prof_r02_fused_${TAG:-sparse}.ncu-rep --page raw --csv > gpurun_out/prof_r02_fused_${TAG:-sparse}_raw.csv 2>/dev/null
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/fprof_plain.json") if l.startswith("{")][-1])
print("pruned", d["pruned"]); print("nosparse", d.get("pruned_without_essential_evaluation")); print("exh", d["exhaustive"])
PY
