#!/bin/bash
# full ncu capture (with source counters) of one traversal launch on config 5 and config 2
mkdir -p gpurun_out
for w in config5 config2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 8 -c 1 \
    -o gpurun_out/r02n_${w} python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02n_${w}.log 2>&1
done
ls -la gpurun_out/
