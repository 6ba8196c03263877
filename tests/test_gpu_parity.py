"""GPU parity: the CUDA path through the C ABI against the fp64 oracle, per iteration.

Tolerance (north_star: "within a stated relative tolerance, e.g. 1e-4 in fp32"): norm-wise per
action node, |gpu - oracle|_inf <= TOL * |oracle|_inf(node) + 1e-6 * |oracle|_inf(table), TOL = 1e-4
(tests/util.py:compare_tables), checked after EVERY iteration: the first iterations of a free run
from zero tables, then lock-step iterations restarted from the oracle's state (util.lockstep).
"""
import numpy as np
import pytest

import rustsolver_b200 as rb
from oracle import OracleGame
from rustsolver_b200 import configs
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _pair(options, card_abs=None, keys=None, **kw):
    n, tree = rb.build_game_tree(options)
    ranges = options.ranges()
    eng = rb.Engine(tree, ranges, options.board_mask, card_abs or [], **kw)
    orc = OracleGame(tree, ranges, options.board_mask, keys=keys)
    return tree, eng, orc


def test_river_small_ranges_per_iteration():
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, eng, orc = _pair(o)
    for it in range(5):
        eng.iterate(1)
        orc.iterate(1)
        util.compare_tables(eng, orc, tree, TOL)
    br_g, br_o = eng.best_response(), orc.best_response()
    ev_g, ev_o = eng.average_value(), orc.average_value()
    assert np.allclose(br_g, br_o, rtol=1e-4, atol=1e-4), (br_g, br_o)
    assert np.allclose(ev_g, ev_o, rtol=1e-4, atol=1e-4), (ev_g, ev_o)


def test_river_default_flop_full_ranges():
    o = rb.default_flop()
    tree, eng, orc = _pair(o, [rb.CardAbstraction.ISOMORPHIC()])
    for it in range(3):
        eng.iterate(1)
        orc.iterate(1)
        util.compare_tables(eng, orc, tree, TOL)
    assert eng.stats().updates_per_iteration == 41078


def test_no_graph_matches_graph():
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    n, tree = rb.build_game_tree(o)
    r = o.ranges()
    e1 = rb.Engine(tree, r, o.board_mask)
    e2 = rb.Engine(tree, r, o.board_mask, flags=rb.RS_FLAG_NO_GRAPH)
    e1.iterate(7)
    e2.iterate(7)
    for an in range(tree.n_actions):
        a, b = e1.read_infoset(an), e2.read_infoset(an)  # same engine layout on both sides: compare raw slabs
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])  # bit-reproducible


def test_chain_split_is_bit_identical():
    """Rounds with one or two boards run with split node tasks (TK_TRAV_TERMS, reach-first TK_DOWN); the flag
    RS_FLAG_NO_CHAIN_SPLIT keeps one task per node.  Same arithmetic in the same order: identical tables."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    r = o.ranges()
    e1 = rb.Engine(tree, r, o.board_mask)
    e2 = rb.Engine(tree, r, o.board_mask, flags=rb.RS_FLAG_NO_CHAIN_SPLIT)
    e1.iterate(5)
    e2.iterate(5)
    st = e1.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    for an, b in util.all_slabs(tree, nb):
        x, y = e1.read_infoset(an, b), e2.read_infoset(an, b)
        assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1]), (an, b)
    assert e1.best_response() == e2.best_response()


def test_street_kernel_matches_the_task_kernel():
    """RS_FLAG_STREET_KERNEL walks the final round with the fused street kernel (csrc/street_kernel.cu: one CTA per (board,
    street segment), terminals by list walks) instead of the per-(node, board) dataflow tasks.  Different summation orders,
    same values to fp32 rounding -- on full ranges (several warps per board) and two streets.  Both engines run in
    lock-step with the fp64 oracle at a fifth of the usual tolerance; comparing the two fp32 engines with each other
    directly is ill-posed on boards where most hands tie (regret matching is discontinuous on rows whose regrets are
    rounding noise, util.lockstep)."""
    o = util.small_options("4d5dAs3c", ["random", "random"], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    r = o.ranges()
    e1 = rb.Engine(tree, r, o.board_mask, flags=rb.RS_FLAG_STREET_KERNEL)
    e2 = rb.Engine(tree, r, o.board_mask)
    kinds1 = {k["kind"] for k in e1.profile_iteration()}
    kinds2 = {k["kind"] for k in e2.profile_iteration()}
    assert "street" in kinds1 and "street" not in kinds2, (kinds1, kinds2)
    for eng in (e1, e2):
        orc = OracleGame(tree, r, o.board_mask)
        al = util.RowAligner(eng, orc, tree)
        util.copy_oracle_to_engine(eng, orc, tree, aligner=al)  # profile_iteration above ran one iteration: back to zero tables
        util.lockstep(eng, orc, tree, n_free=1, n_locked=2, tol=2e-5)
    assert np.allclose(e1.best_response(), e2.best_response(), rtol=1e-5, atol=1e-5)
    assert np.allclose(e1.average_value(), e2.average_value(), rtol=1e-5, atol=1e-5)


def test_turn_river_small_ranges():
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, eng, orc = _pair(o)
    util.lockstep(eng, orc, tree, n_free=1, n_locked=4, tol=TOL)
    br_g, br_o = eng.best_response(), orc.best_response()
    assert np.allclose(br_g, br_o, rtol=1e-4, atol=1e-4), (br_g, br_o)


def test_flop_rooted_small_ranges_with_allin_runouts():
    # short stacks force ALLIN terminals on the flop and turn -> run-out chance nodes
    o = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
    tree, eng, orc = _pair(o)
    assert (tree.ttype[tree.type == 1] == 0).any(), "expected ALLIN terminals"
    util.lockstep(eng, orc, tree, n_free=1, n_locked=2, tol=TOL)


def test_bucketed_rows_many_to_one():
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    r = o.ranges()
    k0 = util.bucket_keys_for(None, r, 1, 7, seed=3)
    k1 = util.bucket_keys_for(None, r, 48, 11, seed=4)
    abs_ = [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k0), rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k1)]
    eng = rb.Engine(tree, r, o.board_mask, abs_)
    orc = OracleGame(tree, r, o.board_mask, keys=[k0, k1])
    util.lockstep(eng, orc, tree, n_free=1, n_locked=4, tol=TOL)


@pytest.mark.parametrize("bucketed,n_sizes", [(False, 3), (True, 3), (False, 4), (True, 4)])
def test_wide_nodes(bucketed, n_sizes):
    """Three bet sizes and three raise sizes (config 3's action abstraction): nodes with five actions, the widest the
    register-resident task bodies are specialised for; four sizes: six actions, the generic bodies (task_down_generic /
    task_trav_generic)."""
    bets = [[0.33, 0.66, 1.0]] if n_sizes == 3 else [[0.25, 0.5, 0.75, 1.0]]
    raises = [[2.0, 3.0, 4.0]] if n_sizes == 3 else [[1.5, 2.0, 3.0, 4.0]]
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], bets, raises)
    n, tree = rb.build_game_tree(o)
    widths = np.diff(tree.child_offset)[tree.type == 0]
    assert widths.max() == n_sizes + 2, widths.max()
    r = o.ranges()
    if bucketed:
        k0 = util.bucket_keys_for(None, r, 1, 9, seed=8)
        eng = rb.Engine(tree, r, o.board_mask, [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k0)])
        orc = OracleGame(tree, r, o.board_mask, keys=[k0])
    else:
        eng = rb.Engine(tree, r, o.board_mask)
        orc = OracleGame(tree, r, o.board_mask)
    util.lockstep(eng, orc, tree, n_free=2, n_locked=3, tol=TOL)
    assert np.allclose(eng.best_response(), orc.best_response(), rtol=1e-4, atol=1e-4)
    assert np.allclose(eng.average_value(), orc.average_value(), rtol=1e-4, atol=1e-4)


def test_degenerate_ranges():
    """One hand against three, one of them blocked by the other's cards on some run-outs; and identical one-hand ranges
    that always collide (no compatible deal: every value is 0)."""
    o = util.small_options("4d5dAs3c", ["AhAc", "KhKc,AhKd,7h6h"], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, eng, orc = _pair(o)
    util.lockstep(eng, orc, tree, n_free=2, n_locked=2, tol=TOL)
    assert np.allclose(eng.best_response(), orc.best_response(), rtol=1e-4, atol=1e-4)


def test_single_iteration_from_oracle_state():
    """Restart the GPU from the oracle's state every iteration: isolates one-step error from drift."""
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, eng, orc = _pair(o)
    orc.iterate(20)
    util.lockstep(eng, orc, tree, n_free=0, n_locked=3, tol=2e-5)


def test_discount_sweep():
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, eng, orc = _pair(o)
    util.lockstep(eng, orc, tree, n_free=1, n_locked=1, tol=TOL)
    eng.discount(0.5)
    orc.discount(0.5)
    util.compare_tables(eng, orc, tree, TOL)
    # engine-side schedule of train()'s monitor thread (cfr.rs:248-261): d = p/(p+1) every `interval` iterations
    n, tree2 = rb.build_game_tree(o)
    e2 = rb.Engine(tree2, o.ranges(), o.board_mask, discount_interval=1, discount_cap=1)
    e3 = rb.Engine(tree2, o.ranges(), o.board_mask)
    e2.iterate(1)
    e3.iterate(1)
    for an in range(tree2.n_actions):
        a, b = e2.read_infoset(an), e3.read_infoset(an)
        assert np.allclose(a[0], 0.5 * b[0], rtol=1e-6, atol=0) and np.allclose(a[1], 0.5 * b[1], rtol=1e-6, atol=0)


def _safe_prune_threshold(orc, tree, nb, q=0.35):
    """A threshold at quantile q of the negative regrets, moved into the middle of the widest nearby gap so that no
    regret sits within fp32 rounding of it (the engine restarts from the oracle's state cast to fp32)."""
    vals = np.concatenate([orc.get_slab(an, b)[0].ravel() for an, b in util.all_slabs(tree, nb)])
    neg = np.sort(vals[vals < 0])
    assert neg.size > 50
    i0 = int(q * neg.size)
    window = neg[max(i0 - 500, 0):i0 + 500]
    gaps = np.diff(window)
    j = int(np.argmax(gaps))
    thr = 0.5 * (window[j] + window[j + 1])
    # fp32 rounding moves a regret or the threshold by at most 6e-8 relative: a gap of 1e-6 keeps every cell on its side
    assert gaps[j] > 1e-6 * abs(thr), "no safe gap near the quantile"
    return float(thr), float((vals <= thr).mean())


@pytest.mark.parametrize("bucketed", [False, True])
def test_pruning_matches_oracle(bucketed):
    """rs_set_prune_threshold (cfr.rs:219,352,379-386): regrets at or below the threshold are frozen."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    if bucketed:
        n, tree = rb.build_game_tree(o)
        r = o.ranges()
        k0 = util.bucket_keys_for(None, r, 1, 7, seed=3)
        k1 = util.bucket_keys_for(None, r, 48, 11, seed=4)
        abs_ = [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k0), rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k1)]
        eng = rb.Engine(tree, r, o.board_mask, abs_)
        orc = OracleGame(tree, r, o.board_mask, keys=[k0, k1])
    else:
        tree, eng, orc = _pair(o)
    orc.iterate(6)
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    al = util.RowAligner(eng, orc, tree)
    for _ in range(3):
        thr, frac = _safe_prune_threshold(orc, tree, nb)
        assert 0.05 < frac < 0.6, frac  # the threshold really freezes a good part of the table
        before = {(an, b): orc.get_slab(an, b)[0].copy() for an, b in util.all_slabs(tree, nb)}
        eng.set_prune_threshold(thr)
        orc.set_prune_threshold(thr)
        util.copy_oracle_to_engine(eng, orc, tree, aligner=al)
        eng.iterate(1)
        orc.iterate(1)
        util.compare_tables(eng, orc, tree, TOL, aligner=al)
        for (an, b), r0 in before.items():  # frozen cells are bit-identical on the GPU too
            g = al.read(an, b)[0]
            m = r0 <= thr
            assert np.array_equal(g[m], r0[m].astype(np.float32)), (an, b)
    eng.set_prune_threshold(float("-inf"))
    orc.set_prune_threshold(float("-inf"))
    util.lockstep(eng, orc, tree, n_free=0, n_locked=1, tol=TOL)


def test_average_strategy_dump(tmp_path):
    """rs_dump_average_strategy: headerless LE fp32, action nodes in index order, boards in id order, [row][A]."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, eng, orc = _pair(o)
    eng.iterate(3)
    path = tmp_path / "avg_strategy.dat"
    n = eng.dump_average_strategy(path)
    data = np.fromfile(path, dtype="<f4")
    assert len(data) == n and path.stat().st_size == 4 * n
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    off = 0
    for an, b in util.all_slabs(tree, nb):  # action nodes ascending, boards ascending inside a node
        s = eng.average_strategy(an, b)
        assert np.array_equal(data[off:off + s.size], s.ravel())
        assert np.allclose(s.sum(axis=1), 1.0, atol=1e-5)
        off += s.size
    assert off == n


def test_exploitability_curve_matches_oracle():
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, eng, orc = _pair(o)
    done = 0
    for target in (10, 50, 200):
        eng.iterate(target - done)
        orc.iterate(target - done)
        done = target
        eg = sum(eng.best_response()) / 2
        eo = sum(orc.best_response()) / 2
        # stated bound: exploitability agrees to 2% relative or 0.02 chips (mbb/g at bb=1: 20)
        assert abs(eg - eo) <= max(0.02 * abs(eo), 0.02), (target, eg, eo)
    assert eg < 3.0


def test_batch_of_river_subgames():
    w = configs.config5(n_subgames=6)
    n, tree = rb.build_game_tree(w.options)
    ranges = configs.workload_ranges(w)
    eng = rb.Engine(tree, ranges, 0, w.card_abs, board_masks=w.board_masks)
    eng.iterate(2)
    from oracle import row_alignment
    for s, bm in enumerate(w.board_masks):
        o = OracleGame(tree, ranges, bm)
        o.iterate(2)
        for an in range(tree.n_actions):
            gr, gs = eng.read_infoset(an, s)
            q = int(tree.player[util.node_of(tree, an)])
            idx = row_alignment(eng.card_table(0, q, s), o.rows(0, q, 0), o.n_rows(0, q, 0))
            gr, gs = gr[idx], gs[idx]
            orr, os_ = o.get_slab(an, 0)
            assert gr.shape == orr.shape
            assert np.abs(gr - orr).max() <= TOL * max(np.abs(orr).max(), 1e-12)
            assert np.abs(gs - os_).max() <= TOL * max(np.abs(os_).max(), 1e-12)


def _sample_paths(rng, board_mask, n_paths, n_cards):
    live = [c for c in range(52) if not (board_mask >> c) & 1]
    firsts = rng.choice(live, n_paths, replace=False)
    paths = []
    for f in firsts:
        p = [int(f)]
        while len(p) < n_cards:
            c = int(rng.choice(live))
            if c not in p:
                p.append(c)
        paths.append(p)
    return paths


@pytest.mark.parametrize("board,n_cards,n_paths", [("4d5dAs3c", 1, 1), ("4d5dAs3c", 1, 5), ("4d5dAs", 2, 1), ("4d5dAs", 2, 4)])
def test_sampled_board_iterations_match_oracle(board, n_cards, n_paths):
    """MCCFR-style board sampling (generate_hand, cfr.rs:100-143): the host samples run-outs, the engine traverses
    only those boards.  Both sides restart from the same state (the oracle after two full iterations) for every
    draw, run ONE sampled iteration on the same paths, and must agree on every slab."""
    rounds = n_cards + 1
    stacks = (500, 500) if n_cards == 1 else (60, 60)
    ranges = [util.RANGE_A, util.RANGE_B] if n_cards == 1 else ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"]
    o = util.small_options(board, ranges, [[0.5, 1.0]] * rounds if n_cards == 1 else [[1.0]] * rounds, [[3.0]] * rounds,
                           pot=35 if n_cards == 1 else 40, stacks=stacks)
    tree, eng, orc = _pair(o)
    orc.iterate(2)
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    state = {key: orc.get_slab(*key) for key in util.all_slabs(tree, nb)}
    al = util.RowAligner(eng, orc, tree)
    rng = np.random.RandomState(100 + n_paths)
    for draw in range(3):
        paths = _sample_paths(rng, o.board_mask, n_paths, n_cards)
        for key, (r, s_) in state.items():
            orc.set_slab(key[0], key[1], r, s_)
        util.copy_oracle_to_engine(eng, orc, tree, aligner=al)
        eng.iterate_sampled(paths)
        orc.iterate_sampled(paths)
        util.compare_tables(eng, orc, tree, TOL, aligner=al)
    # sampling every possible first card is the full iteration (importance weight 1), bit for bit on the engine
    if n_cards == 1:
        live = [c for c in range(52) if not (o.board_mask >> c) & 1]
        e1 = rb.Engine(tree, o.ranges(), o.board_mask)
        e2 = rb.Engine(tree, o.ranges(), o.board_mask)
        for _ in range(2):
            e1.iterate(1)
            e2.iterate_sampled([[c] for c in live])
        for key in util.all_slabs(tree, nb):
            x, y = e1.read_infoset(*key), e2.read_infoset(*key)
            assert np.array_equal(x[0], y[0]) and np.array_equal(x[1], y[1])


def test_sampled_iterations_converge():
    """Chance-sampled CFR drives exploitability down (sanity of the importance weights)."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    eng = rb.Engine(tree, o.ranges(), o.board_mask)
    e0 = sum(eng.best_response()) / 2
    rng = np.random.RandomState(9)
    for it in range(600):
        eng.iterate_sampled(_sample_paths(rng, o.board_mask, 4, 1))
    e1 = sum(eng.best_response()) / 2
    assert e1 < 0.25 * e0, (e0, e1)


# ------------------------------------------------------------------------------------------------
# full-size workloads (BASELINE.json configs): what bench.py times is what is checked here
# ------------------------------------------------------------------------------------------------
def _workload_pair(w, **kw):
    n, tree = rb.build_game_tree(w.options)
    ranges = configs.workload_ranges(w)
    eng = rb.Engine(tree, ranges, w.options.board_mask, w.card_abs, **kw)
    return tree, ranges, eng


def _root_round_keys(w, ranges):
    """Bucket keys of the root round for the oracle, computed here from the workload's cluster_arr with the host hand
    indexer (card_abstraction.rs:204-209: canonical index of hole cards + board -> cluster id), independently of the
    engine's plan: [1 board, H] per player.  Blocked hands get a key too; the oracle never reads it."""
    board = [c for c in range(52) if w.options.board_mask >> c & 1]
    ix = rb.HandIndexer([2, len(board)])
    arr = w.card_abs[0].cluster_arr
    per = []
    for q in range(2):
        cards = np.zeros((len(ranges[q]), 2 + len(board)), dtype=np.uint8)
        cards[:, :2] = np.sort(np.asarray(ranges[q], dtype=np.uint8), axis=1)
        cards[:, 2:] = board
        per.append(arr[ix.index_many(cards)].astype(np.uint32)[None, :])
    return per


@pytest.mark.parametrize("flags", [0, rb.RS_FLAG_STREET_KERNEL])
def test_config2_full_size_lockstep(flags):
    """BASELINE config 2 exactly as bench.py runs it: 1128-hand ranges (several compute warps per CTA), 48 river boards,
    K = 500 turn buckets, two streets.  Covers the cross-warp scan totals, the parent_pos scatter of the street roots,
    TK_GATHER over 48 boards and the chain-round split at full size (cfr.rs:502-522)."""
    w = configs.config2()
    tree, ranges, eng = _workload_pair(w, flags=flags)
    keys = [_root_round_keys(w, ranges), None]
    orc = OracleGame(tree, ranges, w.options.board_mask, keys=keys)
    assert eng.stats().updates_per_iteration == orc.updates_per_iter
    util.lockstep(eng, orc, tree, n_free=1, n_locked=2, tol=TOL)
    br_g, br_o = eng.best_response(), orc.best_response()
    assert np.allclose(br_g, br_o, rtol=1e-4, atol=1e-4), (br_g, br_o)
    ev_g, ev_o = eng.average_value(), orc.average_value()
    assert np.allclose(ev_g, ev_o, rtol=1e-4, atol=1e-4), (ev_g, ev_o)


def test_config3_full_size_lockstep():
    """BASELINE config 3: 1326-slot ranges (352-thread CTAs), 76 action nodes, five-action nodes."""
    w = configs.config3()
    tree, ranges, eng = _workload_pair(w)
    orc = OracleGame(tree, ranges, w.options.board_mask)
    util.lockstep(eng, orc, tree, n_free=1, n_locked=2, tol=TOL)
    assert np.allclose(eng.best_response(), orc.best_response(), rtol=1e-4, atol=1e-4)


def _shard_compare(eng, orc, tree, lo1, hi1, per2, n_locked):
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]

    def mine(k, b):
        return k == 0 or (lo1 <= b < hi1 if k == 1 else lo1 * per2 <= b < hi1 * per2)

    rk = {int(tree.an_index[i]): int(tree.round_idx[i]) for i in range(tree.n_nodes) if tree.type[i] == 0}
    slabs = [(an, b) for an, b in util.all_slabs(tree, nb) if mine(rk[an], b)]
    al = util.RowAligner(eng, orc, tree)
    for it in range(1 + n_locked):
        if it > 0:
            for an, b in slabs:
                r, s = orc.get_slab(an, b)
                al.write(an, b, r, s)
        eng.iterate(1)
        orc.iterate(1)
        scales, table, diffs = {}, {"R": 0.0, "S": 0.0}, {}
        for an, b in slabs:
            gr, gs = al.read(an, b)
            orr, os_ = orc.get_slab(an, b)
            for g, o_, nm in ((gr, orr, "R"), (gs, os_, "S")):
                if o_.size:
                    m = float(np.abs(o_).max())
                    scales[(an, nm)] = max(scales.get((an, nm), 0.0), m)
                    table[nm] = max(table[nm], m)
                    diffs[(an, b, nm)] = float(np.abs(g - o_).max())
        for (an, b, nm), d in diffs.items():
            bound = TOL * scales[(an, nm)] + util.ABS_FLOOR * table[nm]
            assert d <= bound, (it, an, b, nm, d, bound)


@pytest.mark.parametrize("flags", [0, rb.RS_FLAG_STREET_KERNEL])
def test_config4_sub_sampled_turn_cards(flags):
    """BASELINE config 4 (flop-rooted, K = 500 flop buckets, 1176-hand ranges, 2352 river boards) on 2 of its 49 turn cards
    with all their river children: the slice a rank of a 25-way board-sharded run owns, walked without peers
    (RS_FLAG_SHARD_ISOLATED) against the oracle restricted to the same window."""
    w = configs.config4()
    n, tree = rb.build_game_tree(w.options)
    ranges = configs.workload_ranges(w)
    world, rank = 25, 3
    eng = rb.Engine(tree, ranges, w.options.board_mask, w.card_abs, rank=rank, world_size=world,
                    flags=flags | rb.RS_FLAG_SHARD_ISOLATED)
    st = eng.stats()
    lo1, hi1 = rank * st.n_boards[1] // world, (rank + 1) * st.n_boards[1] // world
    assert hi1 - lo1 == 2 and st.n_boards_local[2] == 2 * 48
    keys = [_root_round_keys(w, ranges), None, None]
    orc = OracleGame(tree, ranges, w.options.board_mask, keys=keys)
    orc.set_shard(lo1, hi1)
    _shard_compare(eng, orc, tree, lo1, hi1, st.n_boards[2] // st.n_boards[1], n_locked=1)
    assert np.allclose(eng.best_response(), orc.best_response(), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", ["turn", "flop"])
def test_every_shard_of_a_sharded_run_alone(case):
    """Board sharding without a second GPU: each rank of a 2-way (turn-rooted) or 3-way (flop-rooted) split is created in
    turn as an isolated shard and checked against the oracle restricted to the same boards -- local board numbering,
    per-rank task lists, partial chance sums and the replicated root street, everything but the exchange itself."""
    if case == "turn":
        o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
        world = 2
    else:
        o = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
        world = 3
    n, tree = rb.build_game_tree(o)
    ranges = o.ranges()
    for rank in range(world):
        eng = rb.Engine(tree, ranges, o.board_mask, [], rank=rank, world_size=world, flags=rb.RS_FLAG_SHARD_ISOLATED)
        st = eng.stats()
        lo1, hi1 = rank * st.n_boards[1] // world, (rank + 1) * st.n_boards[1] // world
        orc = OracleGame(tree, ranges, o.board_mask)
        orc.set_shard(lo1, hi1)
        per2 = st.n_boards[2] // st.n_boards[1] if st.n_rounds > 2 else 0
        _shard_compare(eng, orc, tree, lo1, hi1, per2, n_locked=1)
        eng.close()


def _xs_case(case):
    if case == "river":
        return util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]]), None
    if case == "turn_river":
        return util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2), None
    if case == "wide":  # six-action nodes: the generic task bodies
        return util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.25, 0.5, 1.0, 2.0]], [[3.0]]), None
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    r = o.ranges()
    return o, [util.bucket_keys_for(None, r, 1, 7, seed=3), util.bucket_keys_for(None, r, 48, 11, seed=4)]


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("case", ["river", "turn_river", "wide", "bucketed"])
def test_sampled_opponent_actions_match_oracle(case, mode):
    """mccfr()'s opponent arm for every hand at once (cfr.rs:466-475, rs_set_opponent_sampling): same counter-based draws on
    both sides, so the tables agree like those of a full traversal.  A draw whose uniform number sits within fp32 rounding
    of a cumulative probability can fall on different actions in fp32 and fp64 (the oracle reports the smallest margin it
    saw): such a seed is retried with the next one; a real defect fails every seed."""
    o, keys = _xs_case(case)
    errors = []
    for seed in (11, 12, 13):
        n, tree = rb.build_game_tree(o)
        ranges = o.ranges()
        abs_ = [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k) for k in keys] if keys else []
        eng = rb.Engine(tree, ranges, o.board_mask, abs_)
        orc = OracleGame(tree, ranges, o.board_mask, keys=keys)
        eng.iterate(2)
        orc.iterate(2)  # a non-uniform strategy before the first draw
        eng.set_opponent_sampling(mode, seed)
        orc.set_opponent_sampling(mode, seed)
        try:
            util.lockstep(eng, orc, tree, n_free=0, n_locked=3, tol=TOL)
        except AssertionError as e:
            errors.append((seed, orc.xs_min_margin(), str(e)[:200]))
            eng.close()
            continue
        # the sampled tables differ from a full traversal's (the mode is really on) and switching it off restores the full one
        full = OracleGame(tree, ranges, o.board_mask, keys=keys)
        full.iterate(2)
        util.copy_oracle_to_engine(eng, full, tree)
        eng.set_opponent_sampling(0)
        eng.iterate(1)
        full.iterate(1)
        util.compare_tables(eng, full, tree, TOL)
        eng.close()
        return
    raise AssertionError(f"every seed failed: {errors}")
