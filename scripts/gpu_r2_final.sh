#!/bin/bash
# End-of-round evidence on ONE GPU: tests, smoke, the driver's bench line, the reference (CPU) arm,
# config 5 / 4 lines, launch list, ncu --set full captures of the shipped kernels.
mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/final_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/final_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/final_smoke.log
echo "== bench"; timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; tail -c 400 gpurun_out/final_bench.json
echo "== reference arm"; timeout 900 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 > gpurun_out/final_reference.json 2> gpurun_out/final_reference.err; echo "rc=$?"; tail -c 500 gpurun_out/final_reference.json
echo "== config 5"; timeout 1200 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err; echo "rc=$?"; tail -2 gpurun_out/final_c5.err
echo "== config 4"; timeout 900 python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/final_c4.json 2> gpurun_out/final_c4.err; echo "rc=$?"; tail -2 gpurun_out/final_c4.err
echo "== launch list"; CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-probe"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ll_ncu.log 2>&1; echo "rc=$?"
echo "== captures"; bash scripts/gpu_r2_prof.sh > gpurun_out/final_prof.log 2>&1; tail -4 gpurun_out/final_prof.log
