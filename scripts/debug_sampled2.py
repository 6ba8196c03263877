import sys
import numpy as np
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from tests import util
o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
n, tree = rb.build_game_tree(o)
live = [c for c in range(52) if not (o.board_mask >> c) & 1]
def run(mode, iters=3):
    eng = rb.Engine(tree, o.ranges(), o.board_mask, flags=rb.RS_FLAG_NO_GRAPH)
    for it in range(iters):
        if mode == 'full': eng.iterate(1)
        else: eng.iterate_sampled([[c] for c in live])
    out = {}
    st = eng.stats(); nb = [st.n_boards[k] for k in range(st.n_rounds)]
    for an, b in util.all_slabs(tree, nb):
        out[(an, b)] = eng.read_infoset(an, b)
    return out
a = run('full'); b = run('all'); c = run('all')
worst = 0; bad = 0
for k in a:
    for x, y in zip(a[k], b[k]):
        d = np.abs(x - y).max() / max(np.abs(x).max(), 1e-9)
        worst = max(worst, d); bad += d > 1e-5
print('full vs sample-all: worst rel', worst, 'bad', bad)
same = all(np.array_equal(b[k][0], c[k][0]) and np.array_equal(b[k][1], c[k][1]) for k in b)
print('sample-all run-to-run bitwise identical:', same)
# single sampled card, twice from scratch
def one(card, reps=1):
    eng = rb.Engine(tree, o.ranges(), o.board_mask, flags=rb.RS_FLAG_NO_GRAPH)
    eng.iterate(3)
    for _ in range(reps): eng.iterate_sampled([[card]])
    st = eng.stats(); nb = [st.n_boards[k] for k in range(st.n_rounds)]
    return {(an, bb): eng.read_infoset(an, bb) for an, bb in util.all_slabs(tree, nb)}
for card in (44, 24, 19):
    x = one(card); y = one(card)
    same = all(np.array_equal(x[k][0], y[k][0]) and np.array_equal(x[k][1], y[k][1]) for k in x)
    print('card', card, 'deterministic:', same)
