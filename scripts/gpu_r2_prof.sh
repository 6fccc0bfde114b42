#!/bin/bash
# ncu --set full captures of the shipped kernels on the headline workload: the second step's launches of
# block_kernel / select_kernel (level 0) and block_kernel / cand_kernel / select_kernel (level 3)
mkdir -p gpurun_out
python scripts/prof_step.py --level 0 > gpurun_out/prof_l0_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"block_kernel|select_kernel" -s 24 -c 24 -f -o gpurun_out/prof_r02_level0 python scripts/prof_step.py --level 0 > gpurun_out/prof_l0_ncu.log 2>&1
echo "level0 rc=$?"; tail -2 gpurun_out/prof_l0_ncu.log
python scripts/prof_step.py --level 3 > gpurun_out/prof_l3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"block_kernel|select_kernel|cand_kernel" -s 32 -c 32 -f -o gpurun_out/prof_r02_level3 python scripts/prof_step.py --level 3 > gpurun_out/prof_l3_ncu.log 2>&1
echo "level3 rc=$?"; tail -2 gpurun_out/prof_l3_ncu.log
ls -la gpurun_out/*.ncu-rep
