#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02b_tests.log
for w in config2 config5 config3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02b_bench_$w.json 2> gpurun_out/r02b_bench_$w.err
done
timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_config4.json 2> gpurun_out/r02b_bench_config4.err
ls -la gpurun_out/ | tail -12
