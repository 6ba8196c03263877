#!/bin/bash
mkdir -p gpurun_out
timeout 55 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02final_config4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02final_launches_bench.log 2>&1
