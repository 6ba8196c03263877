"""Integer tables of the engine (host-only plan compile, no GPU): board table, card table, infoset offsets,
showdown order.  Bit-exact against the oracle / brute force."""
import itertools

import numpy as np
import pytest

import rustsolver_b200 as rb
from oracle import OracleGame
from rustsolver_b200 import configs
from tests import util
from tests.test_poker import _canon


def _plan(options, card_abs=(), **kw):
    n, tree = rb.build_game_tree(options)
    ranges = options.ranges()
    return tree, ranges, rb.Plan(tree, ranges, options.board_mask, list(card_abs), **kw)


def test_stats_match_survey_appendix_b():
    for name, want in (("default_flop_workload", 41078), ("config1", 21620), ("config3", None)):
        w = getattr(configs, name)()
        n, tree = rb.build_game_tree(w.options)
        ranges = configs.workload_ranges(w)
        p = rb.Plan(tree, ranges, w.options.board_mask, w.card_abs)
        st = p.stats()
        if want is not None:
            assert st.updates_per_iteration == want
        assert st.n_combos == 1070190  # 1081 x 990 non-overlapping river pairs (cfr.rs:73-98)


def test_board_table_enumerates_ordered_deal_sequences():
    o = util.small_options("4d5dAs", ["AA,KK", "QQ,JJ"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
    tree, ranges, p = _plan(o)
    st = p.stats()
    assert list(st.n_boards)[:3] == [1, 49, 49 * 48]
    og = OracleGame(tree, ranges, o.board_mask)
    live = [c for c in range(52) if not (o.board_mask >> c & 1)]
    # turn boards: ascending card order (cfr.rs:63-68)
    for i, c in enumerate(live):
        assert p.board_id(1, [c]) == i
        assert og.board_mask(1, i) == o.board_mask | (1 << c)
    rng = np.random.RandomState(5)
    for _ in range(200):
        c1, c2 = rng.choice(live, 2, replace=False)
        b = p.board_id(2, [int(c1), int(c2)])
        assert og.board_mask(2, b) == o.board_mask | (1 << int(c1)) | (1 << int(c2))
    with pytest.raises(rb.EngineError):
        p.board_id(1, [11])  # 4d (card 11) is already on the board


def test_card_table_none_abstraction_matches_oracle_rows():
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, ranges, p = _plan(o)
    og = OracleGame(tree, ranges, o.board_mask)
    for k, nb in ((0, 1), (1, 48)):
        for q in range(2):
            for b in range(nb):
                from oracle import row_alignment
                assert p.num_rows(k, q, b) == og.n_rows(k, q, b)
                # lossless rows are numbered by the engine's board-local hand order: same blocked hands, and a
                # bijection between the two row numberings (row_alignment asserts both)
                idx = row_alignment(p.card_table(k, q, b), og.rows(k, q, b), og.n_rows(k, q, b))
                assert sorted(idx.tolist()) == list(range(og.n_rows(k, q, b)))


def test_card_table_isomorphic_equals_suit_orbit_partition():
    """ISOMORPHIC rows (card_abstraction.rs:186-214) == first-seen dense ids of the brute-force suit-orbit classes."""
    # a two-tone flop with a pair of suits left symmetric: hearts and clubs are interchangeable
    o = util.small_options("AsKs2d", ["random", "random"], [[1.0]] * 3, [[3.0]] * 3)
    n, tree = rb.build_game_tree(o)
    ranges = o.ranges()
    p = rb.Plan(tree, ranges, o.board_mask, [rb.CardAbstraction.ISOMORPHIC()])
    board = [c for c in range(52) if o.board_mask >> c & 1]
    for q in range(2):
        rows = p.card_table(0, q, 0)
        dense, want = {}, []
        for h in ranges[q]:
            key = _canon(list(h) + board, (2, 3))
            want.append(dense.setdefault(key, len(dense)))
        assert rows.tolist() == want
        assert p.num_rows(0, q, 0) == len(dense) < len(ranges[q])  # some hands really merge


def test_card_table_cluster_arr_and_bucket_table_agree():
    """EMD/OCHS path: cluster_arr[canonical index] (card_abstraction.rs:245-251) vs the same buckets given explicitly."""
    w = configs.config1(lossless=False, K=50)
    n, tree = rb.build_game_tree(w.options)
    ranges = w.options.ranges()
    p1 = rb.Plan(tree, ranges, w.options.board_mask, w.card_abs)
    arr = w.card_abs[0].cluster_arr
    ix = rb.HandIndexer([2, 5])
    board = [c for c in range(52) if w.options.board_mask >> c & 1]
    keys = []
    for q in range(2):
        cards = np.zeros((len(ranges[q]), 7), dtype=np.uint8)
        cards[:, :2] = ranges[q]
        cards[:, 2:] = board
        keys.append(arr[ix.index_many(cards)].reshape(1, -1))
    p2 = rb.Plan(tree, ranges, w.options.board_mask, [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=keys)])
    for q in range(2):
        assert np.array_equal(p1.card_table(0, q, 0), p2.card_table(0, q, 0))
        assert p1.num_rows(0, q, 0) == 50


def test_infoset_offsets_follow_readme_layout():
    """infoset_table[round,player][board][action node] -> [row][A] (README.md:45-47): slabs are contiguous per board,
    nodes of the same (round, player) in ActionNode.index order."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, ranges, p = _plan(o)
    nb = [1, 48]
    for k in range(2):
        for q in range(2):
            nodes = sorted(int(tree.an_index[i]) for i in range(tree.n_nodes)
                           if tree.type[i] == 0 and tree.round_idx[i] == k and tree.player[i] == q)
            expect = 0
            for b in range(nb[k]):
                for an in nodes:
                    off, nr, na = p.infoset_offset(an, b)
                    assert off == expect, (k, q, b, an)
                    assert nr == p.num_rows(k, q, b)
                    node = [i for i in range(tree.n_nodes) if tree.type[i] == 0 and tree.an_index[i] == an][0]
                    assert na == tree.child_offset[node + 1] - tree.child_offset[node]
                    expect += ((nr + 3) & ~3) * na  # rows padded to 4 so that every slab is 16-byte aligned


def test_showdown_order_is_the_oracle_strength_order():
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    tree, ranges, p = _plan(o)
    og = OracleGame(tree, ranges, o.board_mask)
    for b in (0, 7, 31, 47):
        for q in range(2):
            order, cls = p.showdown_order(q, b)
            st = og.strengths(q, b)
            live = np.nonzero(st)[0]
            assert sorted(order.tolist()) == live.tolist()
            s = st[order]
            assert np.all(np.diff(s.astype(np.int64)) >= 0)                     # weakest first
            assert np.array_equal(np.diff(cls) != 0, np.diff(s.astype(np.int64)) != 0)  # ties <=> equal class


def test_sharded_plans_partition_the_boards():
    w = configs.config2(K=20)
    n, tree = rb.build_game_tree(w.options)
    ranges = w.options.ranges()
    total = 0
    seen = []
    for r in range(4):
        p = rb.Plan(tree, ranges, w.options.board_mask, w.card_abs, rank=r, world_size=4)
        st = p.stats()
        assert st.n_boards_local[0] == 1 and st.n_boards_local[1] == 12
        total += st.updates_per_iteration
        for b in range(48):
            try:
                p.card_table(1, 0, b)
                seen.append(b)
            except rb.EngineError:
                pass
    assert sorted(seen) == list(range(48))
    full = rb.Plan(tree, ranges, w.options.board_mask, w.card_abs).stats().updates_per_iteration
    root = sum(rb.Plan(tree, ranges, w.options.board_mask, w.card_abs).infoset_offset(an, 0)[1] * 0 for an in [0])  # noqa
    # river slabs are partitioned; the turn slabs are replicated on each of the 4 ranks
    p0 = rb.Plan(tree, ranges, w.options.board_mask, w.card_abs)
    turn_cells = sum(p0.infoset_offset(int(tree.an_index[i]), 0)[1] * p0.infoset_offset(int(tree.an_index[i]), 0)[2]
                     for i in range(tree.n_nodes) if tree.type[i] == 0 and tree.round_idx[i] == 0)
    assert total == full + 3 * turn_cells


def test_malformed_inputs_return_errors():
    o = rb.default_flop()
    n, tree = rb.build_game_tree(o)
    ranges = o.ranges()
    bad = [ranges[0].copy(), ranges[1].copy()]
    bad[0][0] = bad[0][1]  # duplicate combo
    with pytest.raises(rb.EngineError):
        rb.Plan(tree, bad, o.board_mask)
    with pytest.raises(rb.EngineError):
        rb.Plan(tree, ranges, 0b1)  # invalid board mask
    with pytest.raises(rb.EngineError):
        rb.Plan(tree, ranges, o.board_mask, world_size=2, rank=0)  # single river board: nothing to shard
