#!/bin/bash
# scripts/build_variant.sh <suffix> <extra nvcc flags...>  -> rustsolver_b200/libb200cfr_<suffix>.so (tuning experiments)
set -e
SUF=$1; shift
cd "$(dirname "$0")/.."
OUT=build/variant_$SUF; mkdir -p $OUT
for f in kernels.cu street_kernel.cu indexer_kernel.cu abstraction_kernels.cu histogram_kernel.cu engine.cu plan.cpp poker.cpp game.cpp hand_indexer.cpp trainer.cpp host_api.cpp; do
  x="c++"; [[ $f == *.cu ]] && x="cu"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-O3 "$@" -x $x -c rustsolver_b200/csrc/$f -o $OUT/$f.o &
done
wait
nvcc -shared -o rustsolver_b200/libb200cfr_$SUF.so $OUT/*.o -gencode arch=compute_100a,code=sm_100a -cudart static -ldl -lpthread
echo built rustsolver_b200/libb200cfr_$SUF.so
