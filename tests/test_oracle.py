"""The oracle against its own committed goldens and against itself (naive vs fast terminals, literal vs vector)."""
import json
from pathlib import Path

import numpy as np

import rustsolver_b200 as rb
from oracle import OracleGame
from tests import util

GOLDEN = Path(__file__).parent / "golden"


def _small():
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    _, tree = rb.build_game_tree(o)
    return o, tree


def test_oracle_reproduces_golden_trajectory():
    g = json.loads((GOLDEN / "cfr_small_river.json").read_text())
    o, tree = _small()
    og = OracleGame(tree, o.ranges(), o.board_mask, fast_terminals=True)  # golden was written with the naive terminals
    done = 0
    for it in (1, 2, 5):
        og.iterate(it - done)
        done = it
        for an, (r, s) in g[str(it)].items():
            gr, gs = og.get_slab(int(an), 0)
            assert np.allclose(gr, np.array(r), rtol=1e-10, atol=1e-13)
            assert np.allclose(gs, np.array(s), rtol=1e-10, atol=1e-16)
    assert np.allclose(og.best_response(), g["br_after_5"], rtol=1e-9)
    assert np.allclose(og.average_value(), g["ev_after_5"], rtol=1e-9)


def test_naive_and_fast_terminals_agree_on_two_streets():
    o = util.small_options("4d5dAs3c", ["AA,KK,AKs,76s,54s,T9s", "QQ,JJ,AQs,65s,32s,KQo"], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    _, tree = rb.build_game_tree(o)
    a = OracleGame(tree, o.ranges(), o.board_mask, fast_terminals=False)
    b = OracleGame(tree, o.ranges(), o.board_mask, fast_terminals=True)
    a.iterate(2)
    b.iterate(2)
    for an in a.action_nodes:
        k = int(tree.round_idx[a.action_nodes[an]])
        for bd in range(a.n_boards(k)):
            ra, sa = a.get_slab(an, bd)
            rb_, sb = b.get_slab(an, bd)
            assert np.allclose(ra, rb_, rtol=1e-9, atol=1e-15) and np.allclose(sa, sb, rtol=1e-9, atol=1e-18)


def test_average_strategies_are_zero_sum_and_exploitability_falls():
    o, tree = _small()
    og = OracleGame(tree, o.ranges(), o.board_mask)
    og.iterate(1)
    e1 = sum(og.best_response()) / 2
    og.iterate(199)
    ev = og.average_value()
    assert abs(ev[0] + ev[1]) < 1e-9            # +-pot payoffs are zero-sum (cfr.rs:525-556)
    e200 = sum(og.best_response()) / 2
    assert 0 <= e200 < 0.1 * e1


def test_literal_scalar_cfr_converges_to_the_same_region():
    """(a) literal i32 x10000 in-place cfr() vs (b) vector fp64: agreement at convergence level."""
    o, tree = _small()
    vec = OracleGame(tree, o.ranges(), o.board_mask)
    lit = OracleGame(tree, o.ranges(), o.board_mask)
    vec.iterate(150)
    visited, _ = lit.literal_cfr(150)
    assert visited == int(lit.n_combos)
    lit.literal_to_double(10000.0)
    ev, el = sum(vec.best_response()) / 2, sum(lit.best_response()) / 2
    assert el < 4.0 and ev < 4.0
    assert abs(vec.average_value()[0] - lit.average_value()[0]) < 0.5


def test_literal_mccfr_reduces_exploitability():
    o, tree = _small()
    g = OracleGame(tree, o.ranges(), o.board_mask)
    g.literal_to_double(100.0)
    e0 = sum(g.best_response()) / 2
    g.literal_mccfr(200000, n_threads=2, seed=3)
    g.literal_to_double(100.0)
    e1 = sum(g.best_response()) / 2
    assert e1 < 0.5 * e0


def test_chance_sum_switch_documents_the_reference_quirk():
    """cfr.rs:511-521 returns the SUM over deals while passing reach/len down; the oracle's default is the mean."""
    o = util.small_options("4d5dAs3c", ["AA,KK", "QQ,JJ"], [[1.0]] * 2, [[3.0]] * 2)
    _, tree = rb.build_game_tree(o)
    a = OracleGame(tree, o.ranges(), o.board_mask, chance_sum=False)
    b = OracleGame(tree, o.ranges(), o.board_mask, chance_sum=True)
    _, ua = a.literal_cfr(1)
    _, ub = b.literal_cfr(1)
    assert ua != ub


def test_sampled_runouts_are_an_unbiased_estimate_of_the_full_traversal():
    """MCCFR-style board sampling (generate_hand, cfr.rs:100-143) in vector form: one uniformly sampled next card,
    weighted by the number of possible deals.  Averaged over every card it must equal the full traversal exactly,
    and sampling all cards at once IS the full traversal."""
    o = util.small_options("4d5dAs3c", ["AA,KK,AKs,76s", "QQ,JJ,AQs,65s"], [[1.0]] * 2, [[3.0]] * 2)
    _, tree = rb.build_game_tree(o)
    live = [c for c in range(52) if not (o.board_mask >> c) & 1]
    full = OracleGame(tree, o.ranges(), o.board_mask)
    full.iterate(1)
    acc = None
    for c in live:
        g = OracleGame(tree, o.ranges(), o.board_mask)
        g.iterate_sampled([[c]])
        r = g.get_slab(0, 0)[0]
        acc = r if acc is None else acc + r
    assert np.allclose(acc / len(live), full.get_slab(0, 0)[0], rtol=1e-12, atol=1e-15)
    a = OracleGame(tree, o.ranges(), o.board_mask)
    b = OracleGame(tree, o.ranges(), o.board_mask)
    for _ in range(2):
        a.iterate(1)
        b.iterate_sampled([[c] for c in live])
    for an in a.action_nodes:
        k = int(tree.round_idx[a.action_nodes[an]])
        for bd in range(a.n_boards(k)):
            x, y = a.get_slab(an, bd), b.get_slab(an, bd)
            assert np.allclose(x[0], y[0], rtol=1e-12, atol=1e-18) and np.allclose(x[1], y[1], rtol=1e-12, atol=1e-20)


def test_pruning_freezes_regrets_at_or_below_the_threshold():
    """cfr.rs:352,379-386,419: a pruned action is not explored, so its regret is left alone; -inf = no pruning."""
    o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
    n, tree = rb.build_game_tree(o)
    a, b = OracleGame(tree, o.ranges(), o.board_mask), OracleGame(tree, o.ranges(), o.board_mask)
    a.iterate(5)
    b.iterate(5)
    vals = np.concatenate([a.get_slab(an, 0)[0].ravel() for an in range(tree.n_actions)])
    thr = float(np.quantile(vals[vals < 0], 0.4))
    before = [a.get_slab(an, 0)[0].copy() for an in range(tree.n_actions)]
    a.set_prune_threshold(thr)
    b.set_prune_threshold(float("-inf"))
    a.iterate(1)
    b.iterate(1)
    n_frozen = n_moved = 0
    for an in range(tree.n_actions):
        ra, rb_ = a.get_slab(an, 0)[0], b.get_slab(an, 0)[0]
        m = before[an] <= thr
        assert np.array_equal(ra[m], before[an][m])  # frozen exactly
        n_frozen += int(m.sum())
        n_moved += int((rb_[m] != before[an][m]).sum())
    assert n_frozen > 10 and n_moved > 0  # without pruning those cells do move


def _table_delta(og, tree, before):
    out = []
    for an, b in util.all_slabs(tree, [og.n_boards(k) for k in range(og.n_rounds)]):
        r, _ = og.get_slab(an, b)
        out.append((r - before[(an, b)][0]).ravel())
    return np.concatenate(out)


def test_sampled_opponent_actions_are_unbiased():
    """mccfr()'s opponent arm for every hand at once (cfr.rs:466-475, mode 1): the expected regret update of one traversal
    with one sampled action per opponent hand and node equals the update of the full traversal from the same tables."""
    o, tree = _small()
    og = OracleGame(tree, o.ranges(), o.board_mask, fast_terminals=True)
    og.iterate(3)  # a non-uniform current strategy
    slabs = list(util.all_slabs(tree, [og.n_boards(k) for k in range(og.n_rounds)]))
    state = {(an, b): og.get_slab(an, b) for an, b in slabs}

    def restore():
        for (an, b), (r, s) in state.items():
            og.set_slab(an, b, r, s)

    og.traverse_player(0)
    full = _table_delta(og, tree, state)
    n = 400
    acc = np.zeros_like(full)
    one = None
    for seed in range(n):
        restore()
        og.set_opponent_sampling(1, seed + 1)
        og.traverse_player(0)
        d = _table_delta(og, tree, state)
        one = d if one is None else one
        acc += d
    og.set_opponent_sampling(0)
    mean = acc / n
    rel_mean = np.linalg.norm(mean - full) / np.linalg.norm(full)
    rel_one = np.linalg.norm(one - full) / np.linalg.norm(full)
    assert rel_one > 4 * rel_mean, (rel_one, rel_mean)  # a single sample is far off, the mean converges like 1/sqrt(n)
    assert rel_mean < 0.12, rel_mean
    # mode 2 multiplies the kept reach by sigma of the drawn action (cfr.rs:474): biased, as the reference's code is
    restore()
    og.set_opponent_sampling(2, 1)
    og.traverse_player(0)
    two = _table_delta(og, tree, state)
    og.set_opponent_sampling(0)
    assert np.linalg.norm(two) < np.linalg.norm(one)
