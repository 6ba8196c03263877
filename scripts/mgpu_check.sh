#!/bin/bash
# usage: scripts/mgpu_check.sh N   (under gpurun --gpus N): sharded parity at world N (fused exchange) + bench at N
N=$1
for case in turn flop; do
  RS_FUSED=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29651 tests/mgpu_worker.py $case 2>&1 | grep -E "mgpu_worker|FAILED|Error" | head -5
done
for n in $N $((N/2)); do
  timeout 600 python bench.py --gpus $n --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('gpus', d['n_gpus'], 'iter/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), d['roofline']['per_kernel_ms'])"
done
