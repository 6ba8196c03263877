"""Betting rules and tree builder: C++ host mirror == Python oracle == SURVEY goldens (bit-exact)."""
import json
from pathlib import Path

import numpy as np
import pytest

import rustsolver_b200 as rb
from oracle import tree_oracle
from rustsolver_b200 import configs

GOLDEN = Path(__file__).parent / "golden"

# SURVEY.md Appendix A, verbatim structure: (id, kind, fields)
APPENDIX_A = """
0 P ->[1]
1 A 0/P0 [X,B0.5,B1] ->[2, 22, 31]
2 A 1/P1 [X,B0.5,B1] ->[3, 4, 13]
3 S 35/0
4 A 2/P0 [C,F,R3] ->[5, 6, 7]
5 S 69/0
6 U 35/0
7 A 3/P1 [C,F,R3] ->[8, 9, 10]
8 S 137/1
9 U 69/1
10 A 4/P0 [C,F] ->[11, 12]
11 S 375/0
12 U 137/0
13 A 5/P0 [C,F,R3] ->[14, 15, 16]
14 S 105/0
15 U 35/0
16 A 6/P1 [C,F,R3] ->[17, 18, 19]
17 S 245/1
18 U 105/1
19 A 7/P0 [C,F] ->[20, 21]
20 S 1035/0
21 U 245/0
22 A 8/P1 [C,F,R3] ->[23, 24, 25]
23 S 69/1
24 U 35/1
25 A 9/P0 [C,F,R3] ->[26, 27, 28]
26 S 137/0
27 U 69/0
28 A 10/P1 [C,F] ->[29, 30]
29 S 375/1
30 U 137/1
31 A 11/P1 [C,F,R3] ->[32, 33, 34]
32 S 105/1
33 U 35/1
34 A 12/P0 [C,F,R3] ->[35, 36, 37]
35 S 245/0
36 U 105/0
37 A 13/P1 [C,F] ->[38, 39]
38 S 1035/1
39 U 245/1
""".strip().splitlines()


def _oracle_tree(o: rb.Options):
    aa = o.action_abstraction
    return tree_oracle.build_game_tree(o.stack_sizes, o.board_mask, o.starting_pot, aa.bet_sizes, aa.raise_sizes)


def test_default_flop_matches_survey_appendix_a():
    o = rb.default_flop()
    assert o.board_mask == 0x1100000008840  # "4d5dAs3cKs" = cards {6, 11, 15, 44, 48}
    n, tree = rb.build_game_tree(o)
    assert n == 14 and tree.n_nodes == 40
    assert tree.dump() == APPENDIX_A
    n2, nodes = _oracle_tree(o)
    assert n2 == 14 and tree_oracle.dump(nodes) == APPENDIX_A


# SURVEY.md Appendix B
APPENDIX_B = {
    "config1": dict(nodes=22, action_nodes=8, infoset_actions=20, chance=0, showdown=7, fold=6, allin=0),
    "config2": dict(nodes=361, action_nodes=132, infoset_actions=348, chance=11, showdown=107, fold=108, allin=2,
                    action_nodes_per_round={0: 14, 1: 118}, infoset_actions_per_round={0: 38, 1: 310}),
    "config3": dict(nodes=226, action_nodes=76, infoset_actions=224, chance=0, showdown=75, fold=74, allin=0),
    "config4": dict(nodes=1864, action_nodes=706, infoset_actions=1778, chance=84, showdown=501, fold=536, allin=36,
                    action_nodes_per_round={0: 14, 1: 118, 2: 574}, infoset_actions_per_round={0: 38, 1: 310, 2: 1430}),
}


@pytest.mark.parametrize("name", sorted(APPENDIX_B))
def test_config_trees_match_survey_appendix_b_and_oracle(name):
    w = getattr(configs, name)()
    n, tree = rb.build_game_tree(w.options)
    n2, nodes = _oracle_tree(w.options)
    st = tree_oracle.tree_stats(nodes)
    for k, v in APPENDIX_B[name].items():
        assert st[k] == v, (name, k, st[k], v)
    assert n == n2 == st["action_nodes"]
    assert tree.dump() == tree_oracle.dump(nodes)  # bit-exact: ids, players, indices, chip values, action order
    f = tree_oracle.flatten(nodes)
    for key in ("type", "parent", "child_offset", "children", "player", "an_index", "round_idx", "value", "ttype", "last_to_act"):
        assert np.array_equal(np.asarray(f[key], dtype=np.int64), getattr(tree, key).astype(np.int64)), key


def test_golden_tree_fixture_config2():
    """tests/golden/tree_config2.json was written by scripts/make_golden.py from the Python oracle."""
    g = json.loads((GOLDEN / "tree_config2.json").read_text())
    w = configs.config2()
    _, tree = rb.build_game_tree(w.options)
    assert tree.dump() == g["dump"]


def test_truncating_chip_arithmetic():
    # bet 0.5 * 35 = 17.5 -> 17 (f64 as u32 truncation, state.rs:161), pot 52; raise 3 x 17 = 51 -> pot 103
    o = rb.default_flop()
    _, t = rb.build_game_tree(o)
    assert int(t.value[5]) == 69 and int(t.value[8]) == 137 and int(t.value[11]) == 375
    # 3 x 105 = 315 > floor(0.67 * 465) -> all-in 465: pot 1035
    assert int(t.value[20]) == 1035


def test_allin_threshold_breaks_size_loop():
    # bet sizes iterate until the first one above 0.67 * stack (state.rs:136-144): [1.0, 100.0] keeps both
    w = configs.config1()
    _, t = rb.build_game_tree(w.options)
    root = 1
    kinds = [int(t.action_kind[e]) for e in range(t.child_offset[root], t.child_offset[root + 1])]
    assert kinds == [2, 0, 0]  # Check, Bet 1.0, Bet 100.0 (the second one is the all-in)


def test_invalid_board_mask_is_an_error_not_a_crash():
    o = rb.default_flop()
    o.board_mask = 0b11  # two cards: reference panics "invalid board mask" (state.rs:63)
    with pytest.raises(rb.EngineError):
        rb.build_game_tree(o)


def test_missing_round_sizes_is_an_error():
    o = rb.Options(stack_sizes=[500, 500], board_mask=rb.get_card_mask("4d5dAs3c"), starting_pot=35,
                   action_abstraction=rb.ActionAbstraction(bet_sizes=[[0.5]], raise_sizes=[[3.0]]))
    with pytest.raises(rb.EngineError):  # the reference indexes bet_sizes[1] out of bounds and panics
        rb.build_game_tree(o)
