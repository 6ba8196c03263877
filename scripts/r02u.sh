#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 scripts/mgpu_repro.py 3 > gpurun_out/r02u_repro.log 2>&1
echo "rc=$?" >> gpurun_out/r02u_repro.log
