"""Board-sharded engine on >= 2 GPUs (one process per GPU under torchrun, NCCL) against the oracle."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.gpu


def _run_group(cmd, timeout, env):
    """Run torchrun in its own process group and, on a timeout, kill the WHOLE group: workers that outlive the launcher
    would keep the GPUs busy under every later test."""
    import os
    import signal
    p = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=str(ROOT), env=env, start_new_session=True)
    try:
        out, err = p.communicate(timeout=timeout)
    except subprocess.TimeoutExpired:
        os.killpg(p.pid, signal.SIGKILL)
        out, err = p.communicate()
        out += f"\n[test] killed after {timeout} s"
    return subprocess.CompletedProcess(cmd, p.returncode, out, err)


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("case,fused,xs", [("turn", 0, 0), ("flop", 0, 0), ("batch", 0, 0), ("turn", 1, 0), ("flop", 1, 0), ("sampled", 0, 0),
                                           ("sampled", 1, 0), ("turn", 1, 1), ("turn", 0, 2)])
def test_sharded_engine_matches_oracle(case, fused, xs):
    """fused = 1: the traversal kernel exchanges the chance-node sums itself over peer memory (rs_exchange_import);
    fused = 0: two launches with an ncclAllReduce between them.  xs = 1 / 2: sampled opponent actions (rs_set_opponent_sampling)."""
    import os
    n = _n_gpus()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29641", str(ROOT / "tests" / "mgpu_worker.py"), case]
    r = _run_group(cmd, 600, dict(os.environ, RS_FUSED=str(fused), RS_XS=str(xs)))
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
    r = _run_group(cmd, 300, dict(os.environ))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "mgpu_worker timeout world=2: OK" in r.stdout
