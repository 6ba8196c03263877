"""torchrun worker: board-sharded engine (one rank per GPU) vs the fp64 oracle.

    torchrun --nproc-per-node N tests/mgpu_worker.py [turn|flop|batch]

Every rank builds the same game, owns a slice of the first dealt-card level (or of the subgame batch),
iterates in lock-step, and checks ITS OWN slabs against the oracle; replicated root-street slabs are
checked on every rank.  Exit code 0 = parity on all ranks.
"""
import os
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import rustsolver_b200 as rb  # noqa: E402
from oracle import OracleGame  # noqa: E402
from rustsolver_b200 import configs  # noqa: E402
from tests import util  # noqa: E402

TOL = 1e-4


def new_nccl_id(rank, dev):
    """A fresh NCCL unique id, made on rank 0 and broadcast through the live torch process group."""
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.tensor(list(rb.nccl_unique_id()), dtype=torch.uint8, device=dev)
    dist.broadcast(idt, 0)
    return bytes(idt.cpu().tolist())


def check_case(case, rank, world, local, dev, fused):
    """One sharded-engine parity case on the live ranks.  Returns (ok on ALL ranks, worst diff / bound on this rank, message,
    local boards).  Collective: every rank of the process group must call it with the same arguments."""
    nccl_id = new_nccl_id(rank, dev)
    board_masks = None
    if case == "turn":
        o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    elif case == "flop":
        o = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
    else:
        w = configs.config5(n_subgames=max(8, world))
        o = w.options
        board_masks = w.board_masks
    n, tree = rb.build_game_tree(o)
    ranges = configs.workload_ranges(w) if case == "batch" else o.ranges()
    eng = rb.Engine(tree, ranges, o.board_mask, [], board_masks=board_masks, device=local, rank=rank, world_size=world,
                    nccl_id=nccl_id)
    if fused and case != "batch":
        eng.enable_fused_exchange(dist, dev)  # in-kernel exchange over peer memory instead of the NCCL all-reduce
    st = eng.stats()
    n_iters = 3
    ok = True
    msg = ""
    worst = 0.0
    try:
        if case == "batch":
            n_iters = 2  # free run from zero tables; see tests/util.py:lockstep for why not more
            eng.iterate(n_iters)
            lo = rank * len(board_masks) // world
            hi = (rank + 1) * len(board_masks) // world
            for s in range(lo, hi):
                og = OracleGame(tree, ranges, board_masks[s])
                og.iterate(n_iters)
                from oracle import row_alignment
                for an in range(tree.n_actions):
                    gr, gs = eng.read_infoset(an, s)
                    q = int(tree.player[util.node_of(tree, an)])
                    idx = row_alignment(eng.card_table(0, q, s), og.rows(0, q, 0), og.n_rows(0, q, 0))
                    gr, gs = gr[idx], gs[idx]
                    orr, os_ = og.get_slab(an, 0)
                    for g, oarr in ((gr, orr), (gs, os_)):
                        r_ = float(np.abs(g - oarr).max() / (TOL * max(np.abs(oarr).max(), 1e-12)))
                        worst = max(worst, r_)
                        assert r_ <= 1.0, (s, an, r_)
        else:
            og = OracleGame(tree, ranges, o.board_mask)
            xs = int(os.environ.get("RS_XS", "0"))
            if xs:  # sampled opponent actions: the draws are keyed by GLOBAL board ids, so the sharding must not show
                n_iters = 2
                # Lock-step from the first sampled iteration on (the engine starts from the oracle's state after two full
                # iterations): in a free run a row whose regrets are rounding noise plays an arbitrary strategy in any two
                # arithmetics, which a full traversal does not notice (the row is indifferent) but a sampled one does.
                # A draw whose uniform number sits within fp32 rounding of a cumulative probability can also fall on
                # different actions in fp32 and fp64: the lock-step trajectory is the oracle's, so a dry run of the oracle
                # alone finds (identically on every rank) a seed whose draws all keep a safe margin.
                og.iterate(2)
                seed = None
                for cand in range(7, 27):
                    dry = OracleGame(tree, ranges, o.board_mask)
                    dry.iterate(2)
                    dry.set_opponent_sampling(xs, cand)
                    dry.iterate(n_iters)
                    if dry.xs_min_margin() > 5e-6:
                        seed = cand
                        break
                assert seed is not None, "no seed with a safe margin"
                eng.set_opponent_sampling(xs, seed)
                og.set_opponent_sampling(xs, seed)
            nb = [st.n_boards[k] for k in range(st.n_rounds)]
            lo1 = rank * nb[1] // world
            hi1 = (rank + 1) * nb[1] // world
            per2 = nb[2] // nb[1] if st.n_rounds > 2 else 0

            def mine(k, b):
                if k == 0:
                    return True
                if k == 1:
                    return lo1 <= b < hi1
                return lo1 * per2 <= b < hi1 * per2

            rk = {int(tree.an_index[i]): int(tree.round_idx[i]) for i in range(tree.n_nodes) if tree.type[i] == 0}
            al = util.RowAligner(eng, og, tree)
            for it in range(n_iters):
                if it > 0 or xs:  # lock-step: restart from the oracle's state (see tests/util.py)
                    for an, b in util.all_slabs(tree, nb):
                        if mine(rk[an], b):
                            r, s = og.get_slab(an, b)
                            al.write(an, b, r, s)
                eng.iterate(1)
                og.iterate(1)
                scales, diffs, table = {}, {}, {"R": 0.0, "S": 0.0}
                for an, b in util.all_slabs(tree, nb):
                    orr, os_ = og.get_slab(an, b)
                    for oarr, nm in ((orr, "R"), (os_, "S")):
                        if oarr.size:
                            m = float(np.abs(oarr).max())
                            scales[(an, nm)] = max(scales.get((an, nm), 0.0), m)
                            table[nm] = max(table[nm], m)
                    if not mine(rk[an], b):
                        continue
                    gr, gs = al.read(an, b)
                    for g, oarr, nm in ((gr, orr, "R"), (gs, os_, "S")):
                        if oarr.size:
                            diffs[(an, b, nm)] = float(np.abs(g - oarr).max())
                bad = None
                for (an, b, nm), d in diffs.items():
                    bound = TOL * scales[(an, nm)] + util.ABS_FLOOR * table[nm]
                    worst = max(worst, d / bound)
                    if d > bound and bad is None:
                        bad = (it, an, b, nm, d, bound)
                # every rank learns about a failure before the next collective launch (a rank that stopped iterating would
                # leave its peers waiting inside the exchange for ever)
                agree = torch.tensor([0 if bad is None else 1], device=dev)
                dist.all_reduce(agree, op=dist.ReduceOp.MAX)
                assert bad is None, bad
                assert agree.item() == 0, "a peer rank reported a parity failure"
            # best response goes through the same exchange
            br, obr = eng.best_response(), og.best_response()
            assert np.allclose(br, obr, rtol=1e-4, atol=1e-4), (br, obr)
    except AssertionError as e:
        ok = False
        msg = repr(e)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    local_boards = [st.n_boards_local[k] for k in range(st.n_rounds)]
    eng.close()
    return flag.item() == 1, worst, msg, local_boards


def check_sampled(rank, world, local, dev, fused):
    """Sampled-board iterations (rs_iterate_sampled, run-outs from rs_sample_runouts) on a board-sharded engine: every rank
    passes the same paths and walks those of its own slice; lock-step against the oracle's sampled iteration."""
    nccl_id = new_nccl_id(rank, dev)
    o = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
    n, tree = rb.build_game_tree(o)
    ranges = o.ranges()
    eng = rb.Engine(tree, ranges, o.board_mask, [], device=local, rank=rank, world_size=world, nccl_id=nccl_id)
    if fused:
        eng.enable_fused_exchange(dist, dev)
    og = OracleGame(tree, ranges, o.board_mask)
    og.iterate(2)
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    lo1, hi1 = rank * nb[1] // world, (rank + 1) * nb[1] // world
    per2 = nb[2] // nb[1]
    rk = {int(tree.an_index[i]): int(tree.round_idx[i]) for i in range(tree.n_nodes) if tree.type[i] == 0}
    mine = lambda k, b: k == 0 or (lo1 <= b < hi1 if k == 1 else lo1 * per2 <= b < hi1 * per2)
    slabs = [(an, b) for an, b in util.all_slabs(tree, nb) if mine(rk[an], b)]
    al = util.RowAligner(eng, og, tree)
    ok, msg, worst = True, "", 0.0
    try:
        for draw in range(3):
            paths = rb.sample_runouts(100 + draw, o.board_mask, 2, 5 + draw)  # the same on every rank
            for an, b in slabs:
                r, s = og.get_slab(an, b)
                al.write(an, b, r, s)
            eng.iterate_sampled(paths)
            og.iterate_sampled(paths)
            scales, table, diffs = {}, {"R": 0.0, "S": 0.0}, {}
            for an, b in util.all_slabs(tree, nb):
                orr, os_ = og.get_slab(an, b)
                for oarr, nm in ((orr, "R"), (os_, "S")):
                    if oarr.size:
                        m = float(np.abs(oarr).max())
                        scales[(an, nm)] = max(scales.get((an, nm), 0.0), m)
                        table[nm] = max(table[nm], m)
                if mine(rk[an], b):
                    gr, gs = al.read(an, b)
                    for g, oarr, nm in ((gr, orr, "R"), (gs, os_, "S")):
                        if oarr.size:
                            diffs[(an, b, nm)] = float(np.abs(g - oarr).max())
            for (an, b, nm), d in diffs.items():
                bound = TOL * scales[(an, nm)] + util.ABS_FLOOR * table[nm]
                worst = max(worst, d / bound)
                assert d <= bound, (draw, an, b, nm, d, bound)
    except AssertionError as e:
        ok, msg = False, repr(e)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    eng.close()
    return flag.item() == 1, worst, msg, [st.n_boards_local[k] for k in range(st.n_rounds)]


def check_timeout(rank, world, local, dev):
    """A rank that never launches: the others must come back with an error instead of spinning for ever (bounded waits of
    the in-kernel exchange, rs_set_wait_timeout_ms).  Rank `world - 1` skips its rs_iterate."""
    nccl_id = new_nccl_id(rank, dev)
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    eng = rb.Engine(tree, o.ranges(), o.board_mask, [], device=local, rank=rank, world_size=world, nccl_id=nccl_id)
    assert eng.enable_fused_exchange(dist, dev)
    eng.iterate(1)  # everybody: fine
    eng.set_wait_timeout_ms(400)
    ok = True
    msg = ""
    if rank != world - 1:
        import time
        t0 = time.perf_counter()
        try:
            eng.iterate(1)
            ok, msg = False, "rs_iterate returned although a peer never launched"
        except rb.EngineError as e:
            dt = time.perf_counter() - t0
            ok = "gave up waiting" in str(e) and dt < 20.0
            msg = "" if ok else f"unexpected error / duration: {e} after {dt:.1f} s"
        try:
            eng.iterate(1)
            ok, msg = False, "an aborted engine accepted more work"
        except rb.EngineError:
            pass
    dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    eng.close()
    return flag.item() == 1, 0.0, msg, []


def main():
    case = sys.argv[1] if len(sys.argv) > 1 else "turn"
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    if case in ("timeout", "sampled"):
        if case == "timeout":
            all_ok, worst, msg, boards = check_timeout(rank, world, local, dev)
        else:
            all_ok, worst, msg, boards = check_sampled(rank, world, local, dev, os.environ.get("RS_FUSED", "0") == "1")
        if msg:
            print(f"[rank {rank}] FAILED {msg}", flush=True)
        if rank == 0:
            print(f"mgpu_worker {case} world={world}: {'OK' if all_ok else 'FAILED'}", flush=True)
        dist.barrier()
        dist.destroy_process_group()
        sys.exit(0 if all_ok else 1)
    all_ok, worst, msg, boards = check_case(case, rank, world, local, dev, os.environ.get("RS_FUSED", "0") == "1")
    if msg:
        print(f"[rank {rank}] FAILED {msg}", flush=True)
    if rank == 0:
        print(f"mgpu_worker {case} world={world}: {'OK' if all_ok else 'FAILED'} (boards local {boards}, worst diff/bound {worst:.3f})", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if all_ok else 1)


if __name__ == "__main__":
    main()
