#!/bin/bash
# End-of-round evidence on ONE GPU: tests, smoke, the driver's bench line, the reference
# (CPU) arm, ncu launch list + full capture of the traversal kernel.
mkdir -p gpurun_out
echo "== tests"; timeout 1500 python -m pytest tests -m gpu -q --timeout=900 > gpurun_out/final_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/final_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/final_smoke.log
echo "== bench"; timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; tail -c 600 gpurun_out/final_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "rc=$?"; tail -c 700 gpurun_out/final_reference.json
echo "== profile"; bash scripts/gpu_profile.sh > gpurun_out/profile_sh.log 2>&1; tail -3 gpurun_out/profile_sh.log
