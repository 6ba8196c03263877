#!/bin/bash
mkdir -p gpurun_out
( time timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --no-cpu-baseline ) > gpurun_out/r02v_bench2.json 2> gpurun_out/r02v_bench2.err
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -k "turn-1-1 or turn-0-2" 2>&1 | tail -5 > gpurun_out/r02v_tests.log
