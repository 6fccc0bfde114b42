#!/bin/bash
# one ncu --set full capture of block_kernel launches for a tuning build (BB25_LIB honoured)
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:block_kernel -s ${NCU_SKIP:-5} -c ${NCU_COUNT:-5} -f -o gpurun_out/prof_variant $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"
tail -2 gpurun_out/plain2.log | cut -c1-300
