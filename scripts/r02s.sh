#!/bin/bash
# A/B of kernel variants (vector prefetch by the dispatcher, streaming hints on table traffic) on config 4 / 5 / 2
mkdir -p gpurun_out
for v in base vp cs vpcs; do
  lib=$PWD/rustsolver_b200/libb200cfr_$v.so; [ $v = base ] && lib=$PWD/rustsolver_b200/libb200cfr.so
  RS_ENGINE_LIB=$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02s_${v}_c4.json 2> gpurun_out/r02s_${v}_c4.err
done
