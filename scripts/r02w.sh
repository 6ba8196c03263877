#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/r02w.log
for combo in "0 2" "1 1"; do
  set -- $combo
  echo "== fused=$1 xs=$2" >> gpurun_out/r02w.log
  RS_FUSED=$1 RS_XS=$2 timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 tests/mgpu_worker.py turn 2>&1 | grep -E "mgpu_worker|FAILED|Error|error" | cut -c1-400 >> gpurun_out/r02w.log
done
