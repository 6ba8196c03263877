#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config4" 2>&1 | tail -5 > gpurun_out/r02r_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02r_bench_c4.json 2> gpurun_out/r02r_bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 25 -c 1 -o gpurun_out/r02r_config4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02r_config4.log 2>&1
