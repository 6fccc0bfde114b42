#!/bin/bash
# Round 2, call B: fused-rank batch retrieval -- parity tests, then config 5 / 4 bench lines.
mkdir -p gpurun_out
echo "== fused tests"; timeout 2400 python -m pytest tests/test_gpu_fused.py -m gpu -q --timeout=1500 -x ${PYTEST_ARGS} > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?"; tail -40 gpurun_out/b_tests.log
if [ "${SMALL:-1}" = "1" ]; then
echo "== config 5 small"; timeout 900 python bench.py --config 5 --docs 4000000 --queries 2000 --steps 2 --warmup 1 > gpurun_out/b_c5_small.json 2> gpurun_out/b_c5_small.err; echo "rc=$?"; tail -c 2500 gpurun_out/b_c5_small.json; tail -5 gpurun_out/b_c5_small.err
echo "== config 4 small"; timeout 900 python bench.py --config 4 --docs 2000000 --queries 256 --steps 2 --warmup 1 > gpurun_out/b_c4_small.json 2> gpurun_out/b_c4_small.err; echo "rc=$?"; tail -c 2500 gpurun_out/b_c4_small.json; tail -5 gpurun_out/b_c4_small.err
fi
if [ "${FULL:-0}" = "1" ]; then
echo "== config 5 full"; timeout 1500 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err; echo "rc=$?"; tail -c 3000 gpurun_out/b_c5.json; tail -5 gpurun_out/b_c5.err
echo "== config 4 full"; timeout 1500 python bench.py --config 4 --steps 3 --warmup 2 > gpurun_out/b_c4.json 2> gpurun_out/b_c4.err; echo "rc=$?"; tail -c 3000 gpurun_out/b_c4.json; tail -5 gpurun_out/b_c4.err
fi
