"""torchrun helper: the sharded parity cases several times in ONE process (what bench.py does after its timed region)."""
import os, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import torch
import torch.distributed as dist
from tests import mgpu_worker

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bad = 0
for rep in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    for case, fused in (("turn", True), ("flop", True), ("turn", False)):
        ok, wr, msg, boards = mgpu_worker.check_case(case, rank, world, local, dev, fused)
        if rank == 0 or msg:
            print(f"[rank {rank}] rep {rep} {case} fused={fused}: ok={ok} worst={wr:.3f} {msg}", flush=True)
        bad += 0 if ok else 1
        if not ok:
            break
    if bad:
        break
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if bad else 0)
