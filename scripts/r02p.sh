#!/bin/bash
# sampled opponent actions parity; wide/small rule on the latency-bound configs; ncu evidence for config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sampled_opponent or no_graph or chain_split or config3" 2>&1 | tail -15 > gpurun_out/r02p_tests.log
for w in config3 config1; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02p_bench_$w.json 2> gpurun_out/r02p_bench_$w.err
RS_SMALL_ONLY=1 timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-extra > gpurun_out/r02p_bench_${w}_small.json 2> gpurun_out/r02p_bench_${w}_small.err
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02p_c4_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02p_c4_launches_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 25 -c 1 -o gpurun_out/r02p_config4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02p_config4.log 2>&1
ls -la gpurun_out/
