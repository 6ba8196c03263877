"""Device hand-isomorphism indexer (indexer_kernel.cu) against the host indexer: bit-exact (integer work).

The kernel restates hand_indexer_s::get_index as generate_maps / get_cluster call it
(src/solver/card_abstraction.rs:133,147,167,205); the host indexer is pinned by the reference's own known answers
in tests/test_poker.py, the device one is pinned here by equality with it on the same inputs plus the same KAT."""
import numpy as np
import pytest

import rustsolver_b200 as rb
from rustsolver_b200 import configs
from tests import util

pytestmark = pytest.mark.gpu


def _random_hands(rng, n, n_cards):
    # n distinct-card hands: argsort of random keys = a uniform permutation, keep the first n_cards
    return np.argsort(rng.random((n, 52)), axis=1)[:, :n_cards].astype(np.uint8)


@pytest.mark.parametrize("n_board", [3, 4, 5])
def test_device_indexer_is_bit_exact(n_board):
    rng = np.random.default_rng(100 + n_board)
    ix = rb.HandIndexer([2, n_board])
    hands = _random_hands(rng, 200_000, 2 + n_board)
    host = ix.index_many(hands)
    dev = ix.index_many_gpu(hands)
    assert np.array_equal(host, dev)
    assert int(dev.max()) < ix.size(1)
    # suit relabelling does not change the index (same orbit), on the device too
    perm = np.array([2, 0, 3, 1], dtype=np.uint8)
    relabelled = ((hands >> 2) << 2) | perm[hands & 3]
    assert np.array_equal(ix.index_many_gpu(relabelled), dev)


def test_device_indexer_edge_inputs():
    ix = rb.HandIndexer([2, 5])
    assert len(ix.index_many_gpu(np.zeros((0, 7), dtype=np.uint8))) == 0  # empty batch
    one = np.array([[51, 50, 0, 1, 2, 3, 4]], dtype=np.uint8)
    assert np.array_equal(ix.index_many_gpu(one), ix.index_many(one))
    # four suits with the same configuration cannot happen with 2 + 5 cards, three can: two hole cards in one suit,
    # one board card in each other suit and two more in the first
    three = np.array([[0, 4, 8, 12, 1, 2, 3]], dtype=np.uint8)
    assert np.array_equal(ix.index_many_gpu(three), ix.index_many(three))


def test_reference_kat_test_init_iso_turn_on_the_device():
    """card_abstraction.rs:307-330 (12 888 classes; [51,5,..] ~ [50,5,..], [6,5,..] !~ [50,5,..]) through the kernel."""
    ix = rb.HandIndexer([2, 4])
    rows = [(a, b, 0, 1, 2, t) for a in range(3, 52) for b in range(3, a) for t in range(3, 52) if t != a and t != b]
    idx = ix.index_many_gpu(np.array(rows, dtype=np.uint8))
    assert len(np.unique(idx)) == 12888
    k = ix.index_many_gpu(np.array([[51, 5, 0, 1, 2, 3], [50, 5, 0, 1, 2, 3], [6, 5, 0, 1, 2, 3]], dtype=np.uint8))
    assert k[0] == k[1] and k[2] != k[1]


def test_engine_card_tables_equal_the_host_plan():
    """rs_create builds its card tables from device indices, rs_plan_create from host indices: same tables."""
    o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    abs_ = [rb.CardAbstraction.ISOMORPHIC(), rb.CardAbstraction.ISOMORPHIC()]
    eng = rb.Engine(tree, o.ranges(), o.board_mask, abs_)
    plan = rb.Plan(tree, o.ranges(), o.board_mask, abs_)
    st = eng.stats()
    for k in range(st.n_rounds):
        for q in range(2):
            for b in range(st.n_boards[k]):
                assert np.array_equal(eng.card_table(k, q, b), plan.card_table(k, q, b)), (k, q, b)
