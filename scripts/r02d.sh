#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02d_tests.log
ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02d_prof_c5 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_prof_c5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02d_prof_c2 python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_prof_c2.log 2>&1
ls -la gpurun_out/ | tail -5
