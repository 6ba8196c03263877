"""CPU oracle for the CFR hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` legs may
import this package.  rustsolver_b200 never does.  PARITY UNPINNED for CFR outputs (the reference
cannot be built here and ships no CFR fixtures) — see the header of cfr_oracle.c.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path
from typing import List, Optional, Sequence

import numpy as np

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "liborc.so"

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)
f32p = C.POINTER(C.c_float)
VP = C.c_void_p

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(f"{LIB_PATH} missing: run `python -m rustsolver_b200.build` (builds oracle/liborc.so with gcc)")
    lib = C.CDLL(str(LIB_PATH))
    lib.orc_evaluate.restype = C.c_uint32
    lib.orc_evaluate.argtypes = [u8p, C.c_int]
    lib.orc_create.restype = VP
    lib.orc_create.argtypes = [C.c_int, u8p, u32p, u32p, u8p, u32p, u8p, u32p, u8p, u8p, C.c_int, u8p, C.c_int, u8p,
                               C.c_uint64, C.POINTER(u32p)]
    lib.orc_destroy.argtypes = [VP]
    lib.orc_set_options.argtypes = [VP, C.c_int, C.c_int, C.c_int]
    lib.orc_set_prune_threshold.argtypes = [VP, C.c_double]
    lib.orc_set_opponent_sampling.argtypes = [VP, C.c_int, C.c_uint64]
    lib.orc_xs_min_margin.argtypes = [VP]
    lib.orc_xs_min_margin.restype = C.c_double
    lib.orc_iterate.argtypes = [VP, C.c_int]
    lib.orc_traverse_player.argtypes = [VP, C.c_int]
    lib.orc_best_response.argtypes = [VP, f64p]
    lib.orc_average_value.argtypes = [VP, f64p]
    lib.orc_root_cfv.argtypes = [VP, C.c_int, f64p]
    lib.orc_discount.argtypes = [VP, C.c_double]
    lib.orc_n_rounds.restype = C.c_int
    lib.orc_n_rounds.argtypes = [VP]
    lib.orc_n_boards.restype = C.c_int
    lib.orc_n_boards.argtypes = [VP, C.c_int]
    lib.orc_board_mask.restype = C.c_uint64
    lib.orc_board_mask.argtypes = [VP, C.c_int, C.c_int]
    lib.orc_n_combos.restype = C.c_double
    lib.orc_n_combos.argtypes = [VP]
    lib.orc_n_rows.restype = C.c_int
    lib.orc_n_rows.argtypes = [VP, C.c_int, C.c_int, C.c_int]
    lib.orc_rows.argtypes = [VP, C.c_int, C.c_int, C.c_int, i32p]
    lib.orc_updates_per_iter.restype = C.c_uint64
    lib.orc_updates_per_iter.argtypes = [VP]
    lib.orc_get_slab.restype = C.c_int
    lib.orc_get_slab.argtypes = [VP, C.c_int, C.c_int, C.c_int, f64p]
    lib.orc_set_slab.restype = C.c_int
    lib.orc_set_slab.argtypes = [VP, C.c_int, C.c_int, C.c_int, f64p]
    lib.orc_get_islab.restype = C.c_int
    lib.orc_get_islab.argtypes = [VP, C.c_int, C.c_int, C.c_int, i32p]
    lib.orc_strengths.argtypes = [VP, C.c_int, C.c_int, u32p]
    lib.orc_literal_cfr.restype = C.c_long
    lib.orc_literal_cfr.argtypes = [VP, C.c_int, C.c_int, C.c_int, f64p]
    lib.orc_literal_mccfr.restype = C.c_long
    lib.orc_literal_mccfr.argtypes = [VP, C.c_long, C.c_int, C.c_uint64, C.c_long, C.c_long, C.c_long]
    lib.orc_literal_to_double.argtypes = [VP, C.c_double]
    lib.orc_num_threads.restype = C.c_int
    lib.orc_iterate_sampled.argtypes = [VP, C.c_int, i32p, i32p]
    lib.orc_set_shard.argtypes = [VP, C.c_int, C.c_int]
    lib.orc_chance_partials.restype = C.c_int
    lib.orc_chance_partials.argtypes = [VP, C.c_int, C.c_int, C.c_int, f64p, C.c_int]
    _lib = lib
    return lib


def _abs_lib():
    lib = load()
    if not getattr(lib, "_abs_ready", False):
        lib.orc_emd_1d.restype = C.c_float
        lib.orc_emd_1d.argtypes = [f32p, f32p, C.c_int]
        lib.orc_l2_dist.restype = C.c_float
        lib.orc_l2_dist.argtypes = [f32p, f32p, C.c_int]
        lib.orc_kmeans_predict.restype = C.c_double
        lib.orc_kmeans_predict.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, C.c_int, u32p, f32p]
        lib.orc_update_min_dists.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, f32p]
        lib.orc_kmeans_fit_regular.restype = C.c_float
        lib.orc_kmeans_fit_regular.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, C.c_int, C.c_int, u32p]
        lib.orc_kmeans_init_pp.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, u32p]
        lib.orc_kmeans_init_random.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint64, u32p]
        lib.orc_generate_histograms.argtypes = [u8p, C.c_int, C.c_uint64, C.c_int64, C.c_int, C.c_int, C.c_uint64, f32p]
        lib.orc_kmeans_fit_growbatch.argtypes = [f32p, C.c_int64, C.c_int, f32p, C.c_int, C.c_int, C.c_int, C.c_uint64, u32p, u32p, f32p]
        lib._abs_ready = True
    return lib


def emd_1d(p, q) -> float:
    """emd::emd_1d (gen_abstraction/emd.rs:54-113), fp32, reference operation order."""
    a = np.ascontiguousarray(p, dtype=np.float32)
    b = np.ascontiguousarray(q, dtype=np.float32)
    assert a.shape == b.shape and a.ndim == 1
    return float(_abs_lib().orc_emd_1d(a.ctypes.data_as(f32p), b.ctypes.data_as(f32p), len(a)))


def l2_dist(p, q) -> float:
    a = np.ascontiguousarray(p, dtype=np.float32)
    b = np.ascontiguousarray(q, dtype=np.float32)
    return float(_abs_lib().orc_l2_dist(a.ctypes.data_as(f32p), b.ctypes.data_as(f32p), len(a)))


def kmeans_predict(points, centers, kind: int = 0):
    """Kmeans::predict (kmeans.rs:173-211): (cluster[n] u32, min_dist[n] f32, inertia)."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    c = np.ascontiguousarray(centers, dtype=np.float32)
    cl = np.zeros(len(x), dtype=np.uint32)
    md = np.zeros(len(x), dtype=np.float32)
    inertia = _abs_lib().orc_kmeans_predict(x.ctypes.data_as(f32p), len(x), x.shape[1], c.ctypes.data_as(f32p), len(c), kind,
                                           cl.ctypes.data_as(u32p), md.ctypes.data_as(f32p))
    return cl, md, float(inertia)


def kmeans_fit_regular(points, centers, kind: int = 0, rounds: int = 10):
    """Kmeans::fit_regular (kmeans.rs:497-599): (cluster[n], new centers, inertia = mean upper bound)."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    c = np.array(centers, dtype=np.float32, copy=True, order="C")
    cl = np.zeros(len(x), dtype=np.uint32)
    inertia = _abs_lib().orc_kmeans_fit_regular(x.ctypes.data_as(f32p), len(x), x.shape[1], c.ctypes.data_as(f32p), len(c), kind, rounds,
                                               cl.ctypes.data_as(u32p))
    return cl, c, float(inertia)


def generate_histograms(cards7, n_known: int, first_index: int, samples: int, bins: int, seed: int = 1) -> np.ndarray:
    """generate_histograms (gen_abstraction/main.rs:79-159) with the EHS computed exactly; cards7 [count][7] = the un-indexed hands."""
    c = np.ascontiguousarray(cards7, dtype=np.uint8)
    out = np.zeros((len(c), bins), dtype=np.float32)
    _abs_lib().orc_generate_histograms(c.ctypes.data_as(u8p), n_known, first_index, len(c), samples, bins, seed, out.ctypes.data_as(f32p))
    return out


def kmeans_fit_growbatch(points, centers, batch: int, kind: int = 0, seed: int = 1):
    """Kmeans::fit_growbatch (kmeans.rs:336-494, one pass): (batch indices, cluster[batch], new centers, min_change, inertia)."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    c = np.array(centers, dtype=np.float32, copy=True, order="C")
    idx = np.zeros(batch, dtype=np.uint32)
    cl = np.zeros(batch, dtype=np.uint32)
    stats = np.zeros(2, dtype=np.float32)
    _abs_lib().orc_kmeans_fit_growbatch(x.ctypes.data_as(f32p), len(x), x.shape[1], c.ctypes.data_as(f32p), len(c), kind, batch, seed,
                                        idx.ctypes.data_as(u32p), cl.ctypes.data_as(u32p), stats.ctypes.data_as(f32p))
    return idx, cl, c, float(stats[0]), float(stats[1])


def kmeans_init_pp(points, k: int, kind: int = 0, seed: int = 1) -> np.ndarray:
    """Kmeans::init_pp (kmeans.rs:60-90) with the stated splitmix64 stream: indices of the k points taken as centres."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros(k, dtype=np.uint32)
    _abs_lib().orc_kmeans_init_pp(x.ctypes.data_as(f32p), len(x), x.shape[1], k, kind, seed, out.ctypes.data_as(u32p))
    return out


def kmeans_init_random(points, k: int, n_restarts: int, kind: int = 0, seed: int = 1) -> np.ndarray:
    """Kmeans::init_random (kmeans.rs:103-166): indices of the centres of the most spread out of n_restarts random sets."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    out = np.zeros(k, dtype=np.uint32)
    _abs_lib().orc_kmeans_init_random(x.ctypes.data_as(f32p), len(x), x.shape[1], k, n_restarts, kind, seed, out.ctypes.data_as(u32p))
    return out


def update_min_dists(points, new_center, min_dists, kind: int = 0):
    """update_min_dists (kmeans.rs:603-619), in place on a copy that is returned."""
    x = np.ascontiguousarray(points, dtype=np.float32)
    c = np.ascontiguousarray(new_center, dtype=np.float32)
    md = np.array(min_dists, dtype=np.float32, copy=True)
    _abs_lib().orc_update_min_dists(x.ctypes.data_as(f32p), len(x), x.shape[1], c.ctypes.data_as(f32p), kind, md.ctypes.data_as(f32p))
    return md


def row_alignment(engine_rows_of_slot: np.ndarray, oracle_rows_of_slot: np.ndarray, n_rows: int) -> np.ndarray:
    """The engine numbers the rows of a lossless table by its board-local hand order, the oracle by hand slot.
    Both expose hand slot -> row (card table); this returns idx with engine_row = idx[oracle_row]."""
    idx = np.full(n_rows, -1, dtype=np.int64)
    g = np.asarray(engine_rows_of_slot, dtype=np.int64)
    o = np.asarray(oracle_rows_of_slot, dtype=np.int64)
    live = o >= 0
    assert np.array_equal(live, g != 0xFFFF), "engine and oracle disagree on which hands the board removes"
    idx[o[live]] = g[live]
    assert (idx >= 0).all()
    # same partition of the hands into rows on both sides
    back = np.full(n_rows, -1, dtype=np.int64)
    back[g[live]] = o[live]
    assert np.array_equal(back[idx], np.arange(n_rows)), "engine and oracle rows are not the same partition"
    return idx


def evaluate(cards: Sequence[int]) -> int:
    a = np.asarray(cards, dtype=np.uint8)
    return int(load().orc_evaluate(a.ctypes.data_as(u8p), len(a)))


class OracleGame:
    """Wraps orc_game.  `tree` is any object with the rs_tree arrays as attributes (numpy or lists)."""

    def __init__(self, tree, ranges: Sequence[np.ndarray], board_mask: int,
                 keys: Optional[List[List[Optional[np.ndarray]]]] = None, fast_terminals: bool = True,
                 chance_sum: bool = False):
        self.lib = load()
        g = lambda name, dt: np.ascontiguousarray(getattr(tree, name) if not isinstance(tree, dict) else tree[name], dtype=dt)
        self.type = g("type", np.uint8)
        self.child_offset = g("child_offset", np.uint32)
        self.children = g("children", np.uint32)
        self.player = g("player", np.uint8)
        self.an_index = g("an_index", np.uint32)
        self.round_idx = g("round_idx", np.uint8)
        self.value = g("value", np.uint32)
        self.ttype = g("ttype", np.uint8)
        self.last_to_act = g("last_to_act", np.uint8)
        self.ranges = [np.ascontiguousarray(r, dtype=np.uint8) for r in ranges]
        self._keys_keep = []
        karr = (u32p * 6)()
        if keys is not None:
            for k, per_round in enumerate(keys):
                if per_round is None:
                    continue
                for q in range(2):
                    if per_round[q] is None:
                        continue
                    a = np.ascontiguousarray(per_round[q], dtype=np.uint32)
                    self._keys_keep.append(a)
                    karr[k * 2 + q] = a.ctypes.data_as(u32p)
        p = lambda a, t: a.ctypes.data_as(t)
        self.h = self.lib.orc_create(len(self.type), p(self.type, u8p), p(self.child_offset, u32p), p(self.children, u32p),
                                     p(self.player, u8p), p(self.an_index, u32p), p(self.round_idx, u8p), p(self.value, u32p),
                                     p(self.ttype, u8p), p(self.last_to_act, u8p), len(self.ranges[0]), p(self.ranges[0], u8p),
                                     len(self.ranges[1]), p(self.ranges[1], u8p), board_mask, karr)
        self.lib.orc_set_options(self.h, int(chance_sum), 0, int(fast_terminals))
        self.n_hands = [len(self.ranges[0]), len(self.ranges[1])]
        self.action_nodes = {int(self.an_index[i]): i for i in range(len(self.type)) if self.type[i] == 0}

    def set_options(self, chance_sum=False, fast_terminals=True):
        self.lib.orc_set_options(self.h, int(chance_sum), 0, int(fast_terminals))

    def set_prune_threshold(self, thr: float):
        """Freeze the regrets of actions at or below `thr` (cfr.rs:352,379-386); -inf switches pruning off."""
        self.lib.orc_set_prune_threshold(self.h, float(thr))

    # --- vector-form fp64 CFR ---
    def set_opponent_sampling(self, mode: int, seed: int = 0):
        """mccfr()'s opponent arm for every hand at once (cfr.rs:466-475); the draw is the hash rs_set_opponent_sampling states."""
        self.lib.orc_set_opponent_sampling(self.h, int(mode), int(seed))

    def xs_min_margin(self) -> float:
        """smallest |u - cumulative sigma| of any draw with non-zero reach since set_opponent_sampling"""
        return float(self.lib.orc_xs_min_margin(self.h))

    def iterate(self, n: int = 1):
        self.lib.orc_iterate(self.h, n)

    def board_id_of(self, dealt) -> list:
        """board ids per round (1..) of a run-out given as dealt cards in deal order (ascending-card board table)."""
        ids, mask, bid = [], self.board_mask(0, 0), 0
        for k, c in enumerate(dealt, start=1):
            per = 52 - bin(mask).count("1")
            idx = sum(1 for x in range(int(c)) if not (mask >> x) & 1)
            bid = bid * per + idx
            mask |= 1 << int(c)
            ids.append(bid)
        return ids

    def iterate_sampled(self, paths):
        """One MCCFR-style iteration on sampled run-outs (paths = [[turn, river], ...] dealt cards)."""
        paths = [list(p) for p in paths]
        ids = [self.board_id_of(p) for p in paths]
        r1 = np.ascontiguousarray([i[0] for i in ids], dtype=np.int32)
        r2 = np.ascontiguousarray([i[1] if len(i) > 1 else 0 for i in ids], dtype=np.int32)
        self.lib.orc_iterate_sampled(self.h, len(paths), r1.ctypes.data_as(i32p), r2.ctypes.data_as(i32p))

    def traverse_player(self, p: int):
        self.lib.orc_traverse_player(self.h, p)

    def best_response(self):
        out = (C.c_double * 2)()
        self.lib.orc_best_response(self.h, out)
        return [out[0], out[1]]

    def average_value(self):
        out = (C.c_double * 2)()
        self.lib.orc_average_value(self.h, out)
        return [out[0], out[1]]

    def root_cfv(self, p: int) -> np.ndarray:
        out = np.zeros(self.n_hands[p], dtype=np.float64)
        self.lib.orc_root_cfv(self.h, p, out.ctypes.data_as(f64p))
        return out

    def chance_partials(self, p: int, lo: int, hi: int, cap_nodes: int = 64) -> np.ndarray:
        """Partial values of the round-0 chance nodes over dealt boards [lo, hi) -> [n_chance, H[p]]."""
        out = np.zeros((cap_nodes, self.n_hands[p]), dtype=np.float64)
        n = self.lib.orc_chance_partials(self.h, p, lo, hi, out.ctypes.data_as(f64p), cap_nodes)
        return out[:n].copy()

    def set_shard(self, lo: int, hi: int):
        """Every later traversal deals only the first-level boards [lo, hi) at the round-0 chance nodes (hi <= 0: all)."""
        self.lib.orc_set_shard(self.h, int(lo), int(hi))

    def discount(self, d: float):
        self.lib.orc_discount(self.h, d)

    # --- shapes ---
    @property
    def n_rounds(self) -> int:
        return self.lib.orc_n_rounds(self.h)

    def n_boards(self, k: int) -> int:
        return self.lib.orc_n_boards(self.h, k)

    def board_mask(self, k: int, b: int) -> int:
        return int(self.lib.orc_board_mask(self.h, k, b))

    @property
    def n_combos(self) -> float:
        return self.lib.orc_n_combos(self.h)

    def n_rows(self, k: int, q: int, b: int) -> int:
        return self.lib.orc_n_rows(self.h, k, q, b)

    def rows(self, k: int, q: int, b: int) -> np.ndarray:
        out = np.zeros(self.n_hands[q], dtype=np.int32)
        self.lib.orc_rows(self.h, k, q, b, out.ctypes.data_as(i32p))
        return out

    @property
    def updates_per_iter(self) -> int:
        return int(self.lib.orc_updates_per_iter(self.h))

    def _slab_shape(self, an: int, b: int):
        node = self.action_nodes[an]
        k, q = int(self.round_idx[node]), int(self.player[node])
        A = int(self.child_offset[node + 1] - self.child_offset[node])
        return self.n_rows(k, q, b), A

    def get_slab(self, an: int, b: int = 0):
        """-> (regrets, strategy_sum) float64 [rows, A]."""
        nr, A = self._slab_shape(an, b)
        r = np.zeros((nr, A), dtype=np.float64)
        s = np.zeros((nr, A), dtype=np.float64)
        self.lib.orc_get_slab(self.h, an, b, 0, r.ctypes.data_as(f64p))
        self.lib.orc_get_slab(self.h, an, b, 1, s.ctypes.data_as(f64p))
        return r, s

    def set_slab(self, an: int, b: int, regrets: np.ndarray, ssum: np.ndarray):
        r = np.ascontiguousarray(regrets, dtype=np.float64)
        s = np.ascontiguousarray(ssum, dtype=np.float64)
        self.lib.orc_set_slab(self.h, an, b, 0, r.ctypes.data_as(f64p))
        self.lib.orc_set_slab(self.h, an, b, 1, s.ctypes.data_as(f64p))

    def get_islab(self, an: int, b: int = 0):
        nr, A = self._slab_shape(an, b)
        r = np.zeros((nr, A), dtype=np.int32)
        s = np.zeros((nr, A), dtype=np.int32)
        self.lib.orc_get_islab(self.h, an, b, 0, r.ctypes.data_as(i32p))
        self.lib.orc_get_islab(self.h, an, b, 1, s.ctypes.data_as(i32p))
        return r, s

    def strengths(self, q: int, b: int) -> np.ndarray:
        out = np.zeros(self.n_hands[q], dtype=np.uint32)
        self.lib.orc_strengths(self.h, q, b, out.ctypes.data_as(u32p))
        return out

    # --- literal scalar restatements ---
    def literal_cfr(self, iterations: int = 1, stride: int = 1, offset: int = 0):
        util = (C.c_double * 2)()
        n = self.lib.orc_literal_cfr(self.h, iterations, stride, offset, util)
        return int(n), [util[0], util[1]]

    def literal_mccfr(self, iterations: int, n_threads: int = 8, seed: int = 1, discount_interval: int = 0,
                      discount_cap: int = 0, prune_after: int = -1) -> int:
        return int(self.lib.orc_literal_mccfr(self.h, iterations, n_threads, seed, discount_interval, discount_cap, prune_after))

    def literal_to_double(self, scale: float):
        self.lib.orc_literal_to_double(self.h, scale)

    @property
    def num_threads(self) -> int:
        return self.lib.orc_num_threads()
