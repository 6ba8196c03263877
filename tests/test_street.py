"""Host side of the fused street kernel (csrc/street.h, plan.cpp: TaskGen::build_street), no GPU.

The sorted sweep values every terminal of the final round (cfr.rs:523-558) from one pass over the board's hands in
strength order with running per-card sums.  Here the event stream the plan compiler emits is replayed in numpy
(same arithmetic as street_kernel.cu: sweep_batch, in fp64) for random opponent reach vectors and compared with the
O(H^2) definition: showdown = compatible weaker reach - compatible stronger reach, fold mass = compatible reach.
"""
import numpy as np
import pytest

import rustsolver_b200 as rb
from rustsolver_b200 import configs
from tests import util


def replay_sweep(ev, x, n_read_positions):
    """ev: event words of one board; x: [rows, H_opp] reach by opponent position.  Returns Y [rows, n_read], the
    totals S [rows] and the per-card sums s [rows, 52] exactly as the kernel's sweep leaves them."""
    rows = x.shape[0]
    s = np.zeros((rows, 52))
    S = np.zeros(rows)
    Y = np.zeros((rows, n_read_positions))
    i = 0
    classes = 0
    while i < len(ev):
        nr, na = int(ev[i]) & 0x7FF, (int(ev[i]) >> 11) & 0x7FF
        i += 1
        reads = [int(e) for e in ev[i:i + nr]]
        adds = [int(e) for e in ev[i + nr:i + nr + na]]
        i += nr + na
        classes += 1
        dec = lambda e: (e & 0x7FF, (e >> 11) & 63, (e >> 17) & 63)
        if nr <= 32:
            for e in reads:
                pos, a, b = dec(e)
                Y[:, pos] = S - s[:, a] - s[:, b]
            for e in adds:
                pos, a, b = dec(e)
                s[:, a] += x[:, pos]
                s[:, b] += x[:, pos]
                S += x[:, pos]
            for e in reads:
                pos, a, b = dec(e)
                Y[:, pos] += S - s[:, a] - s[:, b]
        else:  # large class: class-local sums first
            g = np.zeros((rows, 52))
            G = np.zeros(rows)
            for e in adds:
                pos, a, b = dec(e)
                g[:, a] += x[:, pos]
                g[:, b] += x[:, pos]
                G += x[:, pos]
            for e in reads:
                pos, a, b = dec(e)
                Y[:, pos] = 2.0 * (S - s[:, a] - s[:, b]) + (G - g[:, a] - g[:, b])
            s += g
            S += G
    return Y, S, s, classes


def naive_terminal_values(cards_p, str_p, cards_o, str_o, x):
    """x: [rows, n_o] by opponent position.  Returns (showdown [rows, n_p], mass [rows, n_p])."""
    n_p = len(cards_p)
    sd = np.zeros((x.shape[0], n_p))
    mass = np.zeros((x.shape[0], n_p))
    co = np.asarray(cards_o)
    for t in range(n_p):
        a, b = cards_p[t]
        compat = (co[:, 0] != a) & (co[:, 0] != b) & (co[:, 1] != a) & (co[:, 1] != b)
        sign = np.sign(str_p[t].astype(np.int64) - str_o.astype(np.int64))
        sd[:, t] = (x * (compat * sign)[None, :]).sum(axis=1)
        mass[:, t] = (x * compat[None, :]).sum(axis=1)
    return sd, mass


def board_setup(options, board_id=0, card_abs=()):
    n, tree = rb.build_game_tree(options)
    ranges = options.ranges()
    plan = rb.Plan(tree, ranges, options.board_mask, list(card_abs), flags=rb.RS_FLAG_STREET_KERNEL)
    st = plan.stats()
    k = st.n_rounds - 1
    board_cards = None
    return tree, ranges, plan, k


def check_board(plan, ranges, board_cards, board_id, trav, rng, rows=5):
    o = 1 - trav
    ord_p, _ = plan.showdown_order(trav, board_id)
    ord_o, _ = plan.showdown_order(o, board_id)
    cards_p = [tuple(int(c) for c in ranges[trav][s]) for s in ord_p]
    cards_o = [tuple(int(c) for c in ranges[o][s]) for s in ord_o]
    str_p = np.asarray([rb.evaluate(list(c) + board_cards) for c in cards_p], dtype=np.uint32)
    str_o = np.asarray([rb.evaluate(list(c) + board_cards) for c in cards_o], dtype=np.uint32)
    assert (np.diff(str_p.astype(np.int64)) >= 0).all() and (np.diff(str_o.astype(np.int64)) >= 0).all()
    ev = plan.street_events(trav, board_id)
    x = rng.random((rows, len(cards_o)))
    Y, S, s, classes = replay_sweep(ev, x, len(cards_p))
    sd_ref, mass_ref = naive_terminal_values(cards_p, str_p, cards_o, str_o, x)
    ca = np.asarray([c[0] for c in cards_p])
    cb = np.asarray([c[1] for c in cards_p])
    Cm = S[:, None] - s[:, ca] - s[:, cb]
    # identical combo in the opponent's range (it is counted once in S and once in each of its two card sums)
    pos_o = {tuple(sorted(c)): j for j, c in enumerate(cards_o)}
    xs = np.zeros_like(Cm)
    for t, c in enumerate(cards_p):
        j = pos_o.get(tuple(sorted(c)))
        if j is not None:
            xs[:, t] = x[:, j]
    assert np.allclose(Y - Cm, sd_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(sd_ref).max())), "showdown"
    assert np.allclose(Cm + xs, mass_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(mass_ref).max())), "fold mass"
    # every live hand is read once and added once; classes ascend
    n_read = sum(int(e) & 0x7FF for e in _headers(ev))
    n_add = sum((int(e) >> 11) & 0x7FF for e in _headers(ev))
    assert n_read == len(cards_p) and n_add == len(cards_o)
    return classes


def _headers(ev):
    i = 0
    while i < len(ev):
        yield ev[i]
        i += 1 + (int(ev[i]) & 0x7FF) + ((int(ev[i]) >> 11) & 0x7FF)


def _cards_of_mask(mask):
    return [c for c in range(52) if mask >> c & 1]


@pytest.mark.parametrize("trav", [0, 1])
def test_sweep_equals_the_definition_on_asymmetric_ranges(trav):
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, ranges, plan, k = board_setup(o)
    info = plan.street_info(trav)
    assert info["eligible"] == 1 and info["templates"] >= 1, info
    check_board(plan, ranges, _cards_of_mask(o.board_mask), 0, trav, np.random.default_rng(trav))


def test_sweep_full_ranges_with_ties_and_a_board_that_plays():
    # a broadway straight on board: most hands tie (one class far larger than 32 hands -> the class-local path)
    o = util.small_options("AsKdQcJhTs", ["random", "random"], [[1.0]], [[3.0]])
    tree, ranges, plan, k = board_setup(o)
    classes = check_board(plan, ranges, _cards_of_mask(o.board_mask), 0, 0, np.random.default_rng(7), rows=3)
    ev = plan.street_events(0, 0)
    assert max(int(h) & 0x7FF for h in _headers(ev)) > 32
    assert classes < 200


def test_sweep_on_river_boards_below_a_turn_root():
    o = util.small_options("4d5dAs3c", ["random", "random"], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, ranges, plan, k = board_setup(o)
    assert k == 1
    st = plan.stats()
    assert st.n_boards[1] == 48
    rng = np.random.default_rng(3)
    free = [c for c in range(52) if not (o.board_mask >> c & 1)]
    for board_id in (0, 17, 47):
        cards = _cards_of_mask(o.board_mask) + [free[board_id]]
        check_board(plan, ranges, cards, board_id, board_id % 2, rng, rows=2)


def test_street_plan_shapes_of_the_baseline_configs():
    w = configs.config2()
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, flags=rb.RS_FLAG_STREET_KERNEL)
    for p in range(2):
        info = plan.street_info(p)
        # 11 river subtrees + the 2 all-in run-out showdowns, packed into units of at most 32 reach rows
        assert info["eligible"] == 1 and info["segments"] == 13 and info["max_rows"] <= 32 and info["max_batches"] == 1, info
        assert info["down_ops"] + info["up_ops"] >= 118
    w = configs.config3()
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, flags=rb.RS_FLAG_STREET_KERNEL)
    info = plan.street_info(0)
    assert info["eligible"] == 1 and info["segments"] == 1 and info["max_batches"] == 4, info  # 5-action nodes, 113 rows
    w = configs.config1(lossless=False)
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, flags=rb.RS_FLAG_STREET_KERNEL)
    info = plan.street_info(0)
    assert info["eligible"] == 0 and "bucketed" in info["why"], info
    # off unless asked for
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs)
    assert plan.street_info(0)["eligible"] == 0
