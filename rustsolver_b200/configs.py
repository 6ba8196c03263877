"""The BASELINE.json workloads, exactly as SURVEY.md §8(d) / Appendix B define them.

Each builder returns a `Workload`: Options (reference struct), per-round card abstraction, and for
the batched config the list of subgame boards.  Synthetic bucket files stand in for the absent
round_N_{emd,ochs}.dat (the reference's *.dat are git-ignored): buckets are strength quantiles of
the hand on its board, written into a cluster_arr keyed by the canonical (suit-isomorphic) index —
the reference's own file format (card_abstraction.rs:227-229), so the CLUSTER_ARR + indexer path
runs end to end.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .solver import (ActionAbstraction, CardAbstraction, HandIndexer, Options, RS_ABS_BUCKET_TABLE,
                     RS_ABS_CLUSTER_ARR, evaluate, get_card_mask, range_from_string)

BOARD_RIVER = "4d5dAs3cKs"  # options.rs:57 -> cards {6, 11, 15, 44, 48}


@dataclass
class Workload:
    name: str
    options: Options
    card_abs: List[CardAbstraction]
    board_masks: Optional[List[int]] = None  # config 5: one root board per subgame
    pad_ranges_to_1326: bool = False
    notes: str = ""
    expected: dict = field(default_factory=dict)  # SURVEY App. B counts


def _mask_cards(mask: int) -> List[int]:
    return [c for c in range(52) if mask >> c & 1]


def _lowest_cards_mask(mask: int, n: int) -> int:
    m = 0
    for c in _mask_cards(mask)[:n]:
        m |= 1 << c
    return m


def splitmix64(state: int):
    state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    z = state
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return state, z ^ (z >> 31)


def strength_quantile_cluster_arr(board_mask: int, ranges, K: int, seed: int) -> np.ndarray:
    """Synthetic round_N_{emd,ochs}.dat: cluster id per canonical index, K strength-quantile buckets.

    Only the canonical indices reachable from `board_mask` are filled (the rest stay 0), which is
    all generate_maps (card_abstraction.rs:75-184) ever reads for this subgame."""
    board = _mask_cards(board_mask)
    ix = HandIndexer([2, len(board)])
    arr = np.zeros(ix.size(1), dtype=np.uint32)
    hands = np.unique(np.concatenate([np.sort(np.asarray(r, dtype=np.uint8), axis=1) for r in ranges]), axis=0)
    cards = np.zeros((len(hands), 2 + len(board)), dtype=np.uint8)
    cards[:, :2] = hands
    cards[:, 2:] = board
    idx = ix.index_many(cards)
    strength = np.asarray([evaluate(list(c)) for c in cards], dtype=np.uint64)
    # deterministic tie-break by a seeded hash of the canonical index
    h = (idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(seed)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.lexsort((h, strength))
    uniq_sorted = []
    seen = set()
    for i in order:
        v = int(idx[i])
        if v not in seen:
            seen.add(v)
            uniq_sorted.append(v)
    n = len(uniq_sorted)
    for pos, v in enumerate(uniq_sorted):
        arr[v] = pos * K // n
    return arr


def default_flop_workload() -> Workload:
    from .solver import default_flop
    return Workload("default_flop", default_flop(), [CardAbstraction.ISOMORPHIC()],
                    expected=dict(nodes=40, action_nodes=14, infoset_actions=38, showdown=13, fold=12, updates=41078))


def config1(lossless: bool = True, K: int = 200) -> Workload:
    """River-only, fixed board, pot-bet + all-in, OHSC-like buckets (or lossless ISOMORPHIC)."""
    o = Options(stack_sizes=[500, 500], board_mask=get_card_mask(BOARD_RIVER), starting_pot=35,
                action_abstraction=ActionAbstraction(bet_sizes=[[1.0, 100.0]], raise_sizes=[[100.0]]))
    if lossless:
        abs_ = [CardAbstraction.ISOMORPHIC()]
    else:
        arr = strength_quantile_cluster_arr(o.board_mask, o.ranges(), K, seed=1)
        abs_ = [CardAbstraction(RS_ABS_CLUSTER_ARR, cluster_arr=arr)]
    return Workload("config1_river" + ("" if lossless else f"_ohsc{K}"), o, abs_,
                    expected=dict(nodes=22, action_nodes=8, infoset_actions=20, showdown=7, fold=6,
                                  updates=21620 if lossless else None))


def config2(K: int = 500) -> Workload:
    """Turn + river: 48 river outcomes, EMD-like turn buckets, 2 bet sizes per street."""
    bm = _lowest_cards_mask(get_card_mask(BOARD_RIVER), 4)
    o = Options(stack_sizes=[500, 500], board_mask=bm, starting_pot=35,
                action_abstraction=ActionAbstraction(bet_sizes=[[0.5, 1.0]] * 2, raise_sizes=[[3.0]] * 2))
    arr = strength_quantile_cluster_arr(bm, o.ranges(), K, seed=2)
    return Workload("config2_turn_river", o, [CardAbstraction(RS_ABS_CLUSTER_ARR, cluster_arr=arr), CardAbstraction.NONE()],
                    expected=dict(nodes=361, action_nodes=132, chance=11, showdown=107, fold=108, allin=2))


def config3() -> Workload:
    """River, unabstracted 1326-slot ranges, 3 bet sizes + 3 raise sizes."""
    o = Options(stack_sizes=[500, 500], board_mask=get_card_mask(BOARD_RIVER), starting_pot=35,
                action_abstraction=ActionAbstraction(bet_sizes=[[0.33, 0.66, 1.0]], raise_sizes=[[2.0, 3.0, 4.0]]))
    return Workload("config3_river_1326", o, [CardAbstraction.NONE()], pad_ranges_to_1326=True,
                    expected=dict(nodes=226, action_nodes=76, infoset_actions=224, showdown=75, fold=74))


def config4(K: int = 500) -> Workload:
    """Flop-rooted, potential-aware-like flop buckets, turn/river boards sharded across GPUs."""
    bm = _lowest_cards_mask(get_card_mask(BOARD_RIVER), 3)
    o = Options(stack_sizes=[500, 500], board_mask=bm, starting_pot=35,
                action_abstraction=ActionAbstraction(bet_sizes=[[0.5, 1.0]] * 3, raise_sizes=[[3.0]] * 3))
    arr = strength_quantile_cluster_arr(bm, o.ranges(), K, seed=4)
    return Workload("config4_flop", o,
                    [CardAbstraction(RS_ABS_CLUSTER_ARR, cluster_arr=arr), CardAbstraction.NONE(), CardAbstraction.NONE()],
                    expected=dict(nodes=1864, action_nodes=706, chance=84))


def config5(n_subgames: int = 512, seed: int = 5) -> Workload:
    """Batch of independent river subgames on random boards (config 1's tree and ranges)."""
    w = config1(lossless=True)
    boards = []
    state = seed
    for _ in range(n_subgames):
        cards = []
        while len(cards) < 5:
            state, z = splitmix64(state)
            c = z % 52
            if c not in cards:
                cards.append(c)
        m = 0
        for c in cards:
            m |= 1 << c
        boards.append(m)
    w.name = f"config5_batch{n_subgames}"
    w.board_masks = boards
    w.card_abs = [CardAbstraction.NONE()]
    w.pad_ranges_to_1326 = True
    w.expected = dict(updates=21620 * n_subgames)
    return w


def workload_ranges(w: Workload):
    """Ranges handed to the engine: filtered by the root board, or the full 1326 slots when padded."""
    if w.pad_ranges_to_1326:
        full = range_from_string("random", 0)
        return [full.copy(), full.copy()]
    return w.options.ranges()
