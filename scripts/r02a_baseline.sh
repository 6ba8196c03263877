#!/bin/bash
# round-2 starting point: GPU tests, bench lines of configs 2/5/4, source-level ncu capture of the traversal kernel
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02a_tests.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_c2.json 2> gpurun_out/r02a_bench_c2.err
python bench.py --workload config5 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02a_bench_c5.json 2> gpurun_out/r02a_bench_c5.err
timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err
ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 8 -c 1 -o gpurun_out/r02a_prof_c2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_prof_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 8 -c 1 -o gpurun_out/r02a_prof_c5 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_prof_c5.log 2>&1
ls -la gpurun_out/
