#!/bin/bash
# board/segment-major execution order of the final round: parity, bench A/B, ncu traffic of config 4
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config4 or shard" 2>&1 | tail -5 > gpurun_out/r02q_tests.log
RS_BOARD_MAJOR=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "config2 or turn_river or flop_rooted or bucketed or sampled" 2>&1 | tail -5 >> gpurun_out/r02q_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02q_bench_c4.json 2> gpurun_out/r02q_bench_c4.err
RS_TASK_MAJOR=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02q_bench_c4_taskmajor.json 2> gpurun_out/r02q_bench_c4_taskmajor.err
RS_BOARD_MAJOR=1 timeout 600 python bench.py --workload config2 --steps 20 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02q_bench_c2_boardmajor.json 2> gpurun_out/r02q_bench_c2_boardmajor.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:task_kernel -s 25 -c 1 -o gpurun_out/r02q_config4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r02q_config4.log 2>&1
