#!/bin/bash
# round-2 final: full GPU test suite, default bench line (config 4 + extras + CPU baselines), ncu evidence for config 5 / 3 / 1
mkdir -p gpurun_out
timeout 480 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02x_tests.log
( time timeout 330 python bench.py ) > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err
for w in config5 config3 config1; do
timeout 90 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 8 -c 1 \
    -o gpurun_out/r02x_${w} python bench.py --workload $w --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02x_${w}.log 2>&1
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1
