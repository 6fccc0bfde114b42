#!/bin/bash
# One ncu --set full capture: scripts/gpu_prof.sh <kernel regex> <out name> <skip> <count> -- <command...>
# (the same command runs plainly first; B200_PROFILING.md)
mkdir -p gpurun_out
KRE=$1; OUT=$2; SKIP=$3; CNT=$4; shift 5
"$@" > gpurun_out/${OUT}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KRE} -s ${SKIP} -c ${CNT} -f -o gpurun_out/${OUT} "$@" > gpurun_out/${OUT}_ncu.log 2>&1
echo "capture rc=$?"; tail -3 gpurun_out/${OUT}_ncu.log
