"""Writes tests/golden/* from the ORACLE side only (Python tree oracle, C oracle).  Committed fixtures pin the
oracle against later edits; regenerate only when the oracle itself is deliberately changed."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rustsolver_b200 as rb  # noqa: E402  (ranges / masks only)
from oracle import OracleGame, evaluate, tree_oracle  # noqa: E402
from rustsolver_b200 import configs  # noqa: E402
from tests import util  # noqa: E402

G = ROOT / "tests" / "golden"
G.mkdir(parents=True, exist_ok=True)

# 1. tree of config 2 from the Python oracle
w = configs.config2()
aa = w.options.action_abstraction
_, nodes = tree_oracle.build_game_tree(w.options.stack_sizes, w.options.board_mask, w.options.starting_pot, aa.bet_sizes, aa.raise_sizes)
(G / "tree_config2.json").write_text(json.dumps({"dump": tree_oracle.dump(nodes), "stats": {k: v for k, v in tree_oracle.tree_stats(nodes).items() if not isinstance(v, dict)}}))

# 2. evaluator known answers: category of hand-picked 7-card hands (oracle brute force)
hands = {
    "royal_flush": "AsKsQsJsTs2h3d", "straight_flush_wheel": "As2s3s4s5sKdKh", "quads": "9s9h9d9cAsKsQs",
    "full_house": "KsKhKd2s2h7c8d", "flush": "As9s7s4s2sKdKh", "straight": "9s8h7d6c5sAsAd", "wheel": "As2h3d4c5sKdQh",
    "trips": "7s7h7dAsKd3c2h", "two_pair": "AsAhKsKd3c3d2h", "pair": "AsAh9s7d5c3d2h", "high_card": "AsQh9s7d5c3d2h",
}
out = {}
for name, s in hands.items():
    cards = [c for c in range(52) if rb.get_card_mask(s) >> c & 1]
    out[name] = {"cards": cards, "category": evaluate(cards) >> 20}
(G / "evaluator_kat.json").write_text(json.dumps(out))

# 3. oracle CFR trajectory on a small river game: regrets/strategy sums after 1, 2, 5 iterations + BR values
o = util.small_options("4d5dAs3cKs", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]], [[3.0]])
_, tree = rb.build_game_tree(o)
og = OracleGame(tree, o.ranges(), o.board_mask, fast_terminals=False)  # naive O(H^2) terminals
snap = {}
done = 0
for it in (1, 2, 5):
    og.iterate(it - done)
    done = it
    snap[str(it)] = {str(an): [x.tolist() for x in og.get_slab(an, 0)] for an in sorted(og.action_nodes)}
snap["br_after_5"] = og.best_response()
snap["ev_after_5"] = og.average_value()
(G / "cfr_small_river.json").write_text(json.dumps(snap))
print("golden fixtures written to", G)
