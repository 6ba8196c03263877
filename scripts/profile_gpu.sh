#!/bin/bash
# usage: scripts/profile_gpu.sh <tag>   (run under gpurun; writes gpurun_out/<tag>_*)
TAG=${1:-r01}
mkdir -p gpurun_out
# launch list: every kernel with its device time (cold-cache, serialised: compare shares)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_launches_bench.log 2>&1
# full capture of the dominant kernel (river-street segment kernel): 2 launches after warm-up
ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 8 -c 2 \
    -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_prof_bench.log 2>&1
ls -la gpurun_out/
