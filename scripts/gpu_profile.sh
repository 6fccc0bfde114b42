#!/bin/bash
# ncu evidence for profiles/: (1) launch list with device times, (2) full capture of the
# traversal kernel.  Each ncu run directly follows a plain run of the same command.
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain1.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:block_kernel -s ${NCU_SKIP:-5} -c ${NCU_COUNT:-5} -f -o gpurun_out/prof_tile $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -3 gpurun_out/plain2.log
ls -la gpurun_out
