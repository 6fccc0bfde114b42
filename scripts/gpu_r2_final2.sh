#!/bin/bash
# End-of-round evidence on ONE GPU, final build: tests, smoke, the driver's bench line, config 5, launch list,
# ncu --set full captures of the shipped kernels (exhaustive block_kernel, pruned block_kernel + group_kernel + cand_kernel,
# fused_group_kernel + fused_block_kernel on a 6 M-document config-5 corpus).
mkdir -p gpurun_out
echo "== tests"; timeout 1800 python -m pytest tests -m gpu -q --timeout=1200 > gpurun_out/final_tests.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/final_tests.log
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/final_smoke.log
echo "== bench"; timeout 1500 python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "rc=$?"; tail -c 300 gpurun_out/final_bench.json
echo "== config 5"; timeout 1200 python bench.py --config 5 --steps 3 --warmup 2 > gpurun_out/final_c5.json 2> gpurun_out/final_c5.err; echo "rc=$?"; tail -2 gpurun_out/final_c5.err
echo "== launch list"; CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-probe"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv $CMD > gpurun_out/final_ll_ncu.log 2>&1; echo "rc=$?"
echo "== captures"
cap() { # name level regex skip count
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$3" -s $4 -c $5 -f -o gpurun_out/prof_r02_$1 python scripts/prof_step.py --level $2 > gpurun_out/prof_$1_ncu.log 2>&1
  echo "$1 rc=$?"
  ncu -i gpurun_out/prof_r02_$1.ncu-rep --page raw --csv > gpurun_out/prof_r02_$1_raw.csv 2>/dev/null
}
python scripts/prof_step.py --level 3 > gpurun_out/prof_plain.log 2>&1; echo "plain rc=$?"
cap level0_group3 0 "block_kernel<.int.8, .bool.0, .bool.0, .int.5" 5 1
cap level3_group 3 "bb25::group_kernel<" 5 1
cap level3_pass 3 "block_kernel<.int.8, .bool.0, .bool.0, .int.4" 5 1
rm -f gpurun_out/prof_r02_level3_pass.ncu-rep
A="--config 5 --docs 6000000 --queries 2000 --steps 1 --warmup 1 --no-cpu"
ncu --set full --clock-control none --import-source on -k regex:"fused_group_kernel|fused_block_kernel" -s 16 -c 2 -f -o gpurun_out/prof_r02_fused_final python bench.py $A > gpurun_out/prof_fused_ncu.log 2>&1; echo "fused rc=$?"
ncu -i gpurun_out/prof_r02_fused_final.ncu-rep --page raw --csv > gpurun_out/prof_r02_fused_final_raw.csv 2>/dev/null
ls -la gpurun_out | head -40
