#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02c_tests.log
for w in config2 config5 config3; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_$w.json 2> gpurun_out/r02c_bench_$w.err
done
RS_SWEEP_WARPS=8 timeout 300 python bench.py --workload config2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_config2_sw8.json 2> gpurun_out/r02c_bench_config2_sw8.err
RS_SWEEP_WARPS=2 timeout 300 python bench.py --workload config2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02c_bench_config2_sw2.json 2> gpurun_out/r02c_bench_config2_sw2.err
timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_bench_config4.json 2> gpurun_out/r02c_bench_config4.err
ls -la gpurun_out/ | tail -12
