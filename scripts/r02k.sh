#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
( time python bench.py --impl reference ) > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02k_tests.log
