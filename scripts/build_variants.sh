#!/bin/bash
# Tuning builds of libbb25 (not shipped): scripts/build_variants.sh "<name>:<nvcc -D flags>" ...
# e.g. scripts/build_variants.sh "c5k1:-DBB25_BLOCK_CTAS=5 -DBB25_PASS_CHUNKS=1" "qc16:-DBB25_QC=16"
cd "$(dirname "$0")/.."
mkdir -p build_variants
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared $flags \
    -o build_variants/libbb25_$name.so bayesian_bm25_b200/csrc/*.cu &
done
wait
ls -la build_variants
