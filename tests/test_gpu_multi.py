"""Board-sharded engine on >= 2 GPUs (one process per GPU under torchrun, NCCL) against the oracle."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case,fused", [("turn", 0), ("flop", 0), ("batch", 0), ("turn", 1), ("flop", 1), ("sampled", 0), ("sampled", 1)])
def test_sharded_engine_matches_oracle(case, fused):
    """fused = 1: the traversal kernel exchanges the chance-node sums itself over peer memory (rs_exchange_import);
    fused = 0: two launches with an ncclAllReduce between them."""
    import os
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(ROOT / "tests" / "mgpu_worker.py"), case]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=dict(os.environ, RS_FUSED=str(fused)))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"mgpu_worker {case} world={world}: OK" in r.stdout


def test_a_rank_that_never_launches_is_an_error_not_a_hang():
    """Bounded waits of the in-kernel exchange (rs_set_wait_timeout_ms): the last rank skips an rs_iterate, the others get
    RS_ERR_CUDA within the bound and the aborted engine refuses further work."""
    import os
    if _n_gpus() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29642", str(ROOT / "tests" / "mgpu_worker.py"), "timeout"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=str(ROOT), env=dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mgpu_worker timeout world=2: OK" in r.stdout
