#!/bin/bash
# launch list (per-kernel device times, cold & serialised: compare shares) of a shard-sized step
mkdir -p gpurun_out
CMD="python bench.py --docs ${DOCS:-1100000} --steps 1 --warmup 1 --no-cpu --no-probe"
$CMD > gpurun_out/ll_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ll_launches.csv $CMD > gpurun_out/ll_ncu.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ll_plain.log | cut -c1-300
