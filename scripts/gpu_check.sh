#!/bin/bash
# One gpurun call: smoke, GPU parity tests, short bench.  Everything under `timeout`
# so a hung kernel cannot hold the box.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
echo "== smoke"; timeout 600 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== tests"; timeout 2400 python -m pytest tests -m gpu -q --timeout=1200 -x ${PYTEST_ARGS} > gpurun_out/tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/tests.log
echo "== bench"; timeout 1500 python bench.py --steps ${BENCH_STEPS:-3} --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.log; tail -5 gpurun_out/bench.err
