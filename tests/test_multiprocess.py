"""N>1 host logic on CPU with world_size-2 gloo: board partition, id broadcast plumbing and the one exchange
step of the path — the all-reduce of the chance-node values — checked with the oracle."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rustsolver_b200 as rb
        from oracle import OracleGame
        from tests import util

        # (1) the 128-byte communicator id travels from rank 0 exactly like bench.py does it
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            ident = torch.arange(128, dtype=torch.uint8)
        dist.broadcast(ident, 0)
        assert ident.tolist() == list(range(128))

        o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
        n, tree = rb.build_game_tree(o)
        ranges = o.ranges()
        # (2) the plan of this rank owns a contiguous slice of the 48 river boards; slices tile the level
        plan = rb.Plan(tree, ranges, o.board_mask, [], rank=rank, world_size=world)
        st = plan.stats()
        nb = st.n_boards[1]
        lo, hi = rank * nb // world, (rank + 1) * nb // world
        assert st.n_boards_local[1] == hi - lo
        owned = []
        for b in range(nb):
            try:
                plan.card_table(1, 0, b)
                owned.append(b)
            except rb.EngineError:
                pass
        assert owned == list(range(lo, hi))
        cnt = torch.tensor([hi - lo], dtype=torch.int64)
        dist.all_reduce(cnt)
        assert cnt.item() == nb

        # (3) partial chance-node values over the local boards, summed across ranks == the unsharded values
        og = OracleGame(tree, ranges, o.board_mask)
        og.iterate(3)  # identical (replicated, deterministic) tables on every rank
        for p in range(2):
            part = torch.from_numpy(og.chance_partials(p, lo, hi))
            dist.all_reduce(part)
            full = og.chance_partials(p, 0, nb)
            assert part.shape == full.shape and part.shape[0] == 11  # the 11 public chance nodes of the turn street
            assert np.allclose(part.numpy(), full, rtol=1e-12, atol=1e-18)
        ret[rank] = "ok"
    except Exception as e:  # pragma: no cover
        ret[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_board_sharding():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: "ok", 1: "ok"}, dict(ret)
