#!/bin/bash
# ncu --set full captures of the shipped kernels on the headline workload (second step of prof_step.py):
#   level 0: the largest block_kernel launch (group 3) and the final select_kernel
#   level 3: cand_kernel + its select_kernel, and the group-3 block_kernel + final select_kernel
# Only a few launches per capture: the merged gpurun_out/ must stay below 64 MiB.
mkdir -p gpurun_out
cap() { # name level skip count
  python scripts/prof_step.py --level $2 > gpurun_out/prof_$1_plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:"block_kernel|select_kernel|cand_kernel" -s $3 -c $4 -f -o gpurun_out/prof_r02_$1 python scripts/prof_step.py --level $2 > gpurun_out/prof_$1_ncu.log 2>&1
  echo "$1 rc=$?"; tail -1 gpurun_out/prof_$1_ncu.log
  ncu -i gpurun_out/prof_r02_$1.ncu-rep --page raw --csv > gpurun_out/prof_r02_$1_raw.csv 2>/dev/null
}
cap level0_group3 0 40 2
cap level3_cand 3 32 2
cap level3_group3 3 56 2
ls -la gpurun_out/
