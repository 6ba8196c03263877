"""Shared helpers for the parity tests: build the same game for the engine and for the oracle."""
from __future__ import annotations

import numpy as np

import rustsolver_b200 as rb
from oracle import OracleGame
from rustsolver_b200 import configs


def small_options(board: str, ranges, bets, raises, pot=35, stacks=(500, 500)) -> rb.Options:
    return rb.Options(stack_sizes=list(stacks), board_mask=rb.get_card_mask(board), starting_pot=pot,
                      hand_ranges=list(ranges),
                      action_abstraction=rb.ActionAbstraction(bet_sizes=bets, raise_sizes=raises))


RANGE_A = "AA,KK,QQ,JJ,TT,99,AKs,AQs,KQs,JTs,T9s,76s,54s,AKo,72o"
RANGE_B = "88+,ATs+,KJs+,QJs,65s,AQo+,32s,A2s"


def bucket_keys_for(plan_or_oracle_boards, ranges, n_boards, K, seed):
    """Deterministic pseudo-random bucket keys [n_boards, H] per player (exercise many-to-one rows)."""
    out = []
    for q in range(2):
        H = len(ranges[q])
        rng = np.random.RandomState(seed * 2 + q)
        out.append(rng.randint(0, K, size=(n_boards, H)).astype(np.uint32))
    return out


def node_of(tree, an):
    m = getattr(tree, "_an_to_node", None)
    if m is None:
        m = {int(tree.an_index[i]): i for i in range(tree.n_nodes) if tree.type[i] == 0}
        tree._an_to_node = m
    return m[an]


class RowAligner:
    """engine row <-> oracle row per (round, player, board), through the two card tables."""

    def __init__(self, engine, oracle, tree):
        self.engine, self.oracle, self.tree = engine, oracle, tree
        self.cache = {}

    def idx(self, an, b):
        from oracle import row_alignment
        node = node_of(self.tree, an)
        k, q = int(self.tree.round_idx[node]), int(self.tree.player[node])
        key = (k, q, b)
        if key not in self.cache:
            self.cache[key] = row_alignment(self.engine.card_table(k, q, b), self.oracle.rows(k, q, b), self.oracle.n_rows(k, q, b))
        return self.cache[key]

    def read(self, an, b):
        """engine slabs with rows re-ordered to the oracle's row numbering"""
        r, s = self.engine.read_infoset(an, b)
        i = self.idx(an, b)
        return r[i], s[i]

    def write(self, an, b, regrets, ssum):
        i = self.idx(an, b)
        r = np.zeros_like(regrets, dtype=np.float32)
        s = np.zeros_like(ssum, dtype=np.float32)
        r[i] = regrets
        s[i] = ssum
        self.engine.write_infoset(an, b, r, s)


def all_slabs(tree, n_boards_of_round):
    """(an_index, board) pairs of every infoset slab."""
    for i in range(tree.n_nodes):
        if tree.type[i] == 0:
            k = int(tree.round_idx[i])
            for b in range(n_boards_of_round[k]):
                yield int(tree.an_index[i]), b


ABS_FLOOR = 1e-6  # of the whole table's max magnitude: fp32 cancellation noise floor


def compare_tables(engine, oracle: OracleGame, tree, tol, boards=None, aligner=None):
    """Norm-wise parity of every infoset slab, regrets and strategy sums.

    For each action node n and board b:
        |gpu - oracle|_inf over slab (n, b) <= tol * scale(n) + ABS_FLOOR * scale(table)
    scale(n) = max over the node's boards of |oracle|_inf, scale(table) = max over every slab of the
    same array.  The absolute term covers slabs whose exact value is a pure cancellation (all actions
    of a row worth the same: fp64 leaves ~1e-21, fp32 ~1e-11 next to regrets of 1e-3).
    Returns the worst ratio diff / bound."""
    st = engine.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    diffs, scales, table = {}, {}, {"regret": 0.0, "ssum": 0.0}
    al = aligner or RowAligner(engine, oracle, tree)
    for an, b in all_slabs(tree, nb):
        if boards is not None and b not in boards:
            continue
        gr, gs = al.read(an, b)
        orr, os_ = oracle.get_slab(an, b)
        assert gr.shape == orr.shape, (an, b, gr.shape, orr.shape)
        for g, o, name in ((gr, orr, "regret"), (gs, os_, "ssum")):
            if o.size == 0:
                continue
            assert np.isfinite(g).all(), (an, b, name)
            m = float(np.abs(o).max())
            scales[(an, name)] = max(scales.get((an, name), 0.0), m)
            table[name] = max(table[name], m)
            diffs[(an, b, name)] = float(np.abs(g.astype(np.float64) - o).max())
    worst, where = 0.0, None
    for (an, b, name), d in diffs.items():
        bound = tol * scales[(an, name)] + ABS_FLOOR * table[name]
        err = d / max(bound, 1e-300)
        if err > worst:
            worst, where = err, (an, b, name, d, bound)
    assert worst <= 1.0, f"parity violated: diff/bound = {worst:.3e} at (an, board, array, diff, bound) = {where}"
    return worst


def lockstep(engine, oracle: OracleGame, tree, n_free, n_locked, tol):
    """n_free iterations run freely from zero tables on both sides, compared after each; then n_locked
    iterations where the engine is first reset to the oracle's state (cast to fp32).  Regret matching is
    discontinuous where a row's positive regret mass is rounding noise, so a free run diverges on such
    rows after a few iterations in ANY two arithmetics; lock-step isolates the one-iteration error."""
    al = RowAligner(engine, oracle, tree)
    for _ in range(n_free):
        engine.iterate(1)
        oracle.iterate(1)
        compare_tables(engine, oracle, tree, tol, aligner=al)
    for _ in range(n_locked):
        copy_oracle_to_engine(engine, oracle, tree, aligner=al)
        engine.iterate(1)
        oracle.iterate(1)
        compare_tables(engine, oracle, tree, tol, aligner=al)


def copy_oracle_to_engine(engine, oracle: OracleGame, tree, aligner=None):
    st = engine.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    al = aligner or RowAligner(engine, oracle, tree)
    for an, b in all_slabs(tree, nb):
        r, s = oracle.get_slab(an, b)
        al.write(an, b, r, s)
