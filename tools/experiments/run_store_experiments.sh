#!/bin/bash
# Builds and runs the store-pattern experiments on the GPU box (nvcc is in the image); output -> gpurun_out/.
#   gpurun -- 'bash tools/experiments/run_store_experiments.sh'
set -e
cd "$(dirname "$0")"
mkdir -p ../../gpurun_out
for f in ${EXPERIMENTS:-pattern_store store_policy bulk_store}; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/$f $f.cu
    echo "== $f" | tee -a ../../gpurun_out/r02_store_experiments.txt
    /tmp/$f 2>&1 | tee -a ../../gpurun_out/r02_store_experiments.txt
done
