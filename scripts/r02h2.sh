#!/bin/bash
# hand records as byte offsets + list boundaries written by the scan: full GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 330 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r02h2_tests.log
timeout 110 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02h2_bench.json 2> gpurun_out/r02h2_bench.err
