#!/bin/bash
# per-round block sizes: parity tests, then bench lines (config 4 with config 2 / 5 extras, config 3), A/B against RS_UNIFORM_BLOCK=1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02o_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_c4.json 2> gpurun_out/r02o_bench_c4.err
RS_UNIFORM_BLOCK=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_bench_c4_uniform.json 2> gpurun_out/r02o_bench_c4_uniform.err
timeout 300 python bench.py --workload config3 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02o_bench_c3.json 2> gpurun_out/r02o_bench_c3.err
RS_UNIFORM_BLOCK=1 timeout 300 python bench.py --workload config3 --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02o_bench_c3_uniform.json 2> gpurun_out/r02o_bench_c3_uniform.err
