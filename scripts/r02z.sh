#!/bin/bash
mkdir -p gpurun_out
( time timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 4 --no-cpu-baseline ) > gpurun_out/r02z_bench4.json 2> gpurun_out/r02z_bench4.err
