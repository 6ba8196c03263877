#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "street_kernel_matches" 2>&1 | grep -E "AssertionError|passed|failed|FAILED" > gpurun_out/r02h_tests.log
RS_ENGINE_FLAGS=4 ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02h_prof_c5 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_prof_c5.log 2>&1
RS_ENGINE_FLAGS=4 ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02h_prof_c2 python bench.py --workload config2 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02h_prof_c2.log 2>&1
ls -la gpurun_out | tail -4
