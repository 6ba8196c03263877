#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "street or config2_full or config4 or batch" 2>&1 | grep -E "AssertionError|passed|failed|FAILED|Error" > gpurun_out/r02g_tests.log
for w in config2 config5; do
RS_ENGINE_FLAGS=4 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02g_bench_${w}_street.json 2> gpurun_out/r02g_bench_${w}_street.err
done
RS_ENGINE_FLAGS=4 timeout 600 python bench.py --workload config4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_bench_config4_street.json 2> gpurun_out/r02g_bench_config4_street.err
RS_ENGINE_FLAGS=4 ncu --set full --clock-control none --import-source on -k regex:street_kernel -s 6 -c 1 -o gpurun_out/r02g_prof_c5 python bench.py --workload config5 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_prof_c5.log 2>&1
