#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02m_tests.log
for w in config5 config3 config1; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02m_bench_$w.json 2> gpurun_out/r02m_bench_$w.err
done
