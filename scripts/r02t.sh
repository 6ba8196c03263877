#!/bin/bash
# 2 GPUs: sharded parity tests (incl. sampled opponent actions, the exchange sequence number) and the bench at N = 2 with in-bench parity
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02t_gpus.log
timeout 1500 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15 > gpurun_out/r02t_tests.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --no-cpu-baseline ) > gpurun_out/r02t_bench2.json 2> gpurun_out/r02t_bench2.err
