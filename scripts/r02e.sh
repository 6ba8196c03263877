#!/bin/bash
# round-2 re-entry baseline: GPU tests, bench lines of configs 2/5/3 with the task kernel and with the street kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/r02e_tests.log
for w in config2 config5 config3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_$w.json 2> gpurun_out/r02e_bench_$w.err
done
for w in config2 config5; do
RS_ENGINE_FLAGS=4 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02e_bench_${w}_street.json 2> gpurun_out/r02e_bench_${w}_street.err
done
timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02e_bench_config4.json 2> gpurun_out/r02e_bench_config4.err
ls -la gpurun_out/ | tail -12
