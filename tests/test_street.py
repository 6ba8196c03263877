"""Host side of the fused street kernel (csrc/street.h, plan.cpp: TaskGen::build_street), no GPU.

The street kernel values every terminal of the final round (cfr.rs:523-558) by list walks: one thread per (four reach
rows, list) keeps the running sum of its list -- the opponent hands holding one card, or a piece of the global strength
order -- and follows a precomputed program of the board.  Here the programs the plan compiler emits are replayed in
numpy (same steps as street_kernel.cu: walk_list / walk_chunk / copy-out, in fp64) for random opponent reach vectors
and compared with the O(H^2) definition: showdown = compatible weaker reach - compatible stronger reach, fold mass =
compatible reach.
"""
import numpy as np
import pytest

import rustsolver_b200 as rb
from rustsolver_b200 import configs
from tests import util

START, END = 1 << 23, 1 << 24


def _walk(words, X, Y, rmw_off=None):
    """one piece: returns its total; emits to Y (plain store, or the chunk pass's combine when rmw_off is given)"""
    rows = X.shape[0]
    g, g0, m = np.zeros(rows), np.zeros(rows), np.zeros(rows)
    for w in words:
        w = int(w)
        if w & START:
            g0 = g.copy()
        g = g + X[:, w & 0x7FF]
        if w & END:
            m = g0 + g
        e = (w >> 11) & 0xFFF
        if rmw_off is None:
            Y[:, e] = m
        else:
            assert e < rmw_off
            Y[:, e] = m - Y[:, e] - Y[:, rmw_off + e]
    return g


def replay_programs(prog, x):
    """prog: Plan.street_program(...); x: [rows, n_live_opp] reach by opponent position.  Returns (VY, VM) [rows, HpP]:
    the showdown and mass vectors exactly as the kernel's copy-out leaves them."""
    rows, n_o = x.shape
    HpP, HoP = prog["HpP"], prog["HoP"]
    X = np.zeros((rows, HoP + 1))
    X[:, :n_o] = x
    Y = np.full((rows, 2 * (HpP + 1)), np.nan)  # every live cell must be written before it is read
    Y[:, HpP] = 0.0
    Y[:, 2 * HpP + 1] = 0.0
    LB = np.zeros((rows, 52, 5))  # exclusive prefix of the piece totals of every card list; [4] = the list total
    for c in range(52):
        for j in range(4):
            LB[:, c, j + 1] = LB[:, c, j] + _walk(prog["lists"][:, 4 * c + j], X, Y)
    CB = np.zeros((rows, 129))
    for ch in range(128):
        CB[:, ch + 1] = CB[:, ch] + _walk(prog["chunks"][:, ch], X, Y, rmw_off=HpP + 1)
    tot = CB[:, 128]
    w0, w1 = prog["hinfo"][:, 0].astype(np.int64), prog["hinfo"][:, 1].astype(np.int64)
    c0, c1 = w0 & 63, (w0 >> 6) & 63
    p0lo, p0hi, p1lo, p1hi = (w0 >> 12) & 3, (w0 >> 14) & 7, (w0 >> 17) & 3, (w0 >> 19) & 7
    chlo, chhi, same = w1 & 127, (w1 >> 7) & 255, (w1 >> 15) & 0x7FF
    T = LB[:, :, 4]
    Cp = tot[:, None] - T[:, c0] - T[:, c1]
    VY = (Y[:, :HpP] + CB[:, chlo] + CB[:, chhi] - LB[:, c0, p0lo] - LB[:, c0, p0hi] - LB[:, c1, p1lo] - LB[:, c1, p1hi]) - Cp
    VM = Cp + X[:, same]
    return VY, VM, tot, T


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
    prog = plan.street_program(trav, board_id)
    x = rng.random((rows, len(cards_o)))
    VY, VM, tot, T = replay_programs(prog, x)
    n_p = len(cards_p)
    sd_ref, mass_ref = naive_terminal_values(cards_p, str_p, cards_o, str_o, x)
    assert np.isfinite(VY[:, :n_p]).all(), "a live traverser hand was never emitted to"
    assert np.allclose(VY[:, :n_p], sd_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(sd_ref).max())), "showdown"
    assert np.allclose(VM[:, :n_p], mass_ref, rtol=0, atol=1e-9 * max(1.0, np.abs(mass_ref).max())), "fold mass"
    assert np.allclose(tot, x.sum(axis=1))
    # every opponent hand is added once by the pieces of the global order and once by each of its two card lists; every
    # traverser hand is emitted to once by each of the three
    for words, per_hand in ((prog["lists"], 2), (prog["chunks"], 1)):
        adds = (words.astype(np.int64) & 0x7FF).ravel()
        assert np.array_equal(np.bincount(adds[adds < prog["HoP"]], minlength=len(cards_o))[:len(cards_o)], np.full(len(cards_o), per_hand))
        em = ((words.astype(np.int64) >> 11) & 0xFFF).ravel()
        em = em[em != prog["HpP"]]
        pos = np.where(em > prog["HpP"], em - prog["HpP"] - 1, em)
        assert np.array_equal(np.bincount(pos, minlength=n_p)[:n_p], np.full(n_p, per_hand))
    return prog


def _cards_of_mask(mask):
    return [c for c in range(52) if mask >> c & 1]


@pytest.mark.parametrize("trav", [0, 1])
def test_sweep_equals_the_definition_on_asymmetric_ranges(trav):
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    tree, ranges, plan, k = board_setup(o)
    info = plan.street_info(trav)
    assert info["eligible"] == 1 and info["segments"] >= 1, info
    check_board(plan, ranges, _cards_of_mask(o.board_mask), 0, trav, np.random.default_rng(trav))


def test_sweep_full_ranges_with_ties_and_a_board_that_plays():
    # a broadway straight on board: most hands tie (one class holds most of the range: its piece of the global order is long)
    o = util.small_options("AsKdQcJhTs", ["random", "random"], [[1.0]], [[3.0]])
    tree, ranges, plan, k = board_setup(o)
    prog = check_board(plan, ranges, _cards_of_mask(o.board_mask), 0, 0, np.random.default_rng(7), rows=3)
    assert prog["chunks"].shape[0] <= 32 and prog["lists"].shape[0] <= 64  # the big class is a run of pieces, not one long walk
    w1 = prog["hinfo"][:, 1].astype(np.int64)
    assert ((w1 & 127) != ((w1 >> 7) & 255)).any()


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
        # 11 river subtrees + the 2 all-in run-out showdowns
        assert info["eligible"] == 1 and info["segments"] == 13 and info["max_rows"] <= 40 and info["max_q_sd"] <= 4, info
        assert info["down_ops"] + info["up_ops"] >= 118
        prog = plan.street_program(p, 5)
        assert 8 <= prog["lists"].shape[0] <= 48 and prog["chunks"].shape[0] <= 32, (prog["lists"].shape, prog["chunks"].shape)
    w = configs.config5(n_subgames=4)
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), 0, w.card_abs, board_masks=w.board_masks, flags=rb.RS_FLAG_STREET_KERNEL)
    info = plan.street_info(0)
    assert info["eligible"] == 1 and info["segments"] == 1 and info["max_q_sd"] == 2, info
    w = configs.config3()
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, flags=rb.RS_FLAG_STREET_KERNEL)
    info = plan.street_info(0)
    assert info["eligible"] == 0 and "32 showdown rows" in info["why"], info  # one board, 75 showdowns: the task kernel's job
    w = configs.config1(lossless=False)
    n, tree = rb.build_game_tree(w.options)
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, flags=rb.RS_FLAG_STREET_KERNEL)
    info = plan.street_info(0)
    assert info["eligible"] == 0 and "bucketed" in info["why"], info
    # off unless asked for
    plan = rb.Plan(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs)
    assert plan.street_info(0)["eligible"] == 0
