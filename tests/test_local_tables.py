"""CPU: the board-local index tables the traversal kernel values terminals with (csrc/tasks.h: HandRec, cl_pos entries).

The kernel's scan (kernels.cu: scan_reach) and per-hand terms (hand_terms / hand_mass) are replayed here in numpy on the
tables the plan compiler produced -- prefix sums by position and over the opponent's card lists, the list boundaries
written by the flagged entries with per-thread first ordinals, every gather at the record's ready-made byte offset -- and
compared with the brute-force O(H^2) definition of the two terminal values (cfr.rs:523-558)."""
import numpy as np
import pytest

import rustsolver_b200 as rb
from tests import util

CL_POS_MASK, CL_POS_NONE, CL_FIRST = 0x7FF, 0x7FF, 0x8000


def replay_scan(r, cl, hpad):
    """-> (P, GB, B, total): what scan_reach leaves in shared memory for opponent reach r (by position)."""
    n_threads = (2 * hpad + 7) // 8
    P = np.concatenate([[0.0], np.cumsum(r[:hpad])])
    pos = cl & CL_POS_MASK
    y = np.where(pos != CL_POS_NONE, r[np.minimum(pos, hpad - 1)], 0.0)
    GB = np.concatenate([[0.0], np.cumsum(y)])
    B = np.full(64, np.nan)
    B[54] = B[55] = 0.0
    for t in range(n_threads):
        e = cl[8 * t:8 * t + 8].astype(np.uint32)
        ord0 = int((e[0] >> 11) & 7) | (int((e[1] >> 11) & 7) << 3)
        o = 0 if t == 0 else ord0
        for i in range(8):
            if e[i] & CL_FIRST:
                B[o] = GB[8 * t + i]
                o += 1
        if t == 0:
            B[ord0] = GB[-1]
    return P, GB, B, P[-1]


def replay_terms(rec, r, P, GB, B, total):
    w0, w1, w2, w3 = [int(x) for x in rec]
    k0, k1, same4 = w3 & 0x1FF, (w3 >> 9) & 0x1FF, w3 >> 18
    g0s, g0e, g1s, g1e = B[k0 // 4], B[k0 // 4 + 1], B[k1 // 4], B[k1 // 4 + 1]
    mass = total - (g0e - g0s) - (g1e - g1s) + (r[same4 // 4] if same4 != 0x3FFF else 0.0)
    sd = (P[(w0 & 0xFFFF) // 4] + P[(w0 >> 16) // 4] - total - GB[(w1 & 0xFFFF) // 4] - GB[(w1 >> 16) // 4] + g0s + g0e
          - GB[(w2 & 0xFFFF) // 4] - GB[(w2 >> 16) // 4] + g1s + g1e)
    return mass, sd


@pytest.mark.parametrize("board,ranges", [("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B]), ("4d5dAs3cKs", ["random", "AA,KK,72o,65s"]),
                                          ("2s2h2dAcKd", ["22+,A2s+,K2s+", "random"])])
def test_replayed_scan_and_terms_equal_the_definition(board, ranges):
    o = util.small_options(board, ranges, [[1.0]], [[3.0]])
    n, tree = rb.build_game_tree(o)
    rg = o.ranges()
    plan = rb.Plan(tree, rg, o.board_mask, [])
    rng = np.random.default_rng(1)
    for p in range(2):
        q = 1 - p
        rec, _, sop_p, nl_p = plan.local_tables(0, p, 0)
        _, cl, sop_o, nl_o = plan.local_tables(0, q, 0)
        hpad_o = len(sop_o)
        hands_p, hands_o = np.asarray(rg[p]), np.asarray(rg[q])
        bcards = [c for c in range(52) if (o.board_mask >> c) & 1]
        str_p = np.array([rb.evaluate(bcards + list(hands_p[s])) for s in sop_p[:nl_p]])
        str_o = np.array([rb.evaluate(bcards + list(hands_o[s])) for s in sop_o[:nl_o]])
        assert (np.diff(str_p) >= 0).all() and (np.diff(str_o) >= 0).all()  # positions ascend with strength on the river
        r = np.zeros(hpad_o)
        r[:nl_o] = rng.random(nl_o)
        r[rng.integers(0, nl_o, 5)] = 0.0
        P, GB, B, total = replay_scan(r, cl, hpad_o)
        for i in range(nl_p):
            mass, sd = replay_terms(rec[i], r, P, GB, B, total)
            mine = set(int(c) for c in hands_p[sop_p[i]])
            ok = np.array([not (mine & set(int(c) for c in hands_o[s])) for s in sop_o[:nl_o]])
            want_mass = r[:nl_o][ok].sum()
            want_sd = r[:nl_o][ok & (str_o < str_p[i])].sum() - r[:nl_o][ok & (str_o > str_p[i])].sum()
            assert abs(mass - want_mass) < 1e-9, (p, i, mass, want_mass)
            assert abs(sd - want_sd) < 1e-9, (p, i, sd, want_sd)


def test_list_entries_are_well_formed_on_every_board_of_a_turn_game():
    o = util.small_options("4d5dAs3c", [util.RANGE_A, "random"], [[1.0]] * 2, [[3.0]] * 2)
    n, tree = rb.build_game_tree(o)
    plan = rb.Plan(tree, o.ranges(), o.board_mask, [])
    st = plan.stats()
    for k in range(st.n_rounds):
        for b in range(st.n_boards[k]):
            for q in range(2):
                rec, cl, sop, nl = plan.local_tables(k, q, b)
                pos = cl & CL_POS_MASK
                assert (pos[:2 * nl] < nl).all() and (pos[2 * nl:] == CL_POS_NONE).all()
                assert not (cl[2 * nl:] & CL_FIRST).any()
                n_lists = int(((cl & CL_FIRST) != 0).sum())
                assert (int(cl[0] >> 11) & 7) | ((int(cl[1] >> 11) & 7) << 3) == n_lists
                started = np.concatenate([[0], np.cumsum((cl & CL_FIRST) != 0)])
                for t in range(1, len(cl) // 8):
                    assert (int(cl[8 * t] >> 11) & 7) | ((int(cl[8 * t + 1] >> 11) & 7) << 3) == started[8 * t]
                _, cl_o, _, _ = plan.local_tables(k, 1 - q, b)  # the record's ordinals count the OPPONENT's non-empty lists
                n_lists_o = int(((cl_o & CL_FIRST) != 0).sum())
                for kk in (rec[:nl, 3] & 0x1FF, (rec[:nl, 3] >> 9) & 0x1FF):
                    assert ((kk // 4 < n_lists_o) | (kk == 54 * 4)).all()
