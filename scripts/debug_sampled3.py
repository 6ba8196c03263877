import sys
import numpy as np
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from oracle import OracleGame
from tests import util
o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
n, tree = rb.build_game_tree(o)
live = [c for c in range(52) if not (o.board_mask >> c) & 1]
base = OracleGame(tree, o.ranges(), o.board_mask); base.iterate(3)
eng = rb.Engine(tree, o.ranges(), o.board_mask, flags=rb.RS_FLAG_NO_GRAPH)
st = eng.stats(); nb = [st.n_boards[k] for k in range(st.n_rounds)]
slabs = list(util.all_slabs(tree, nb))
S0 = {k: base.get_slab(*k) for k in slabs}
al = None
for c in live:
    g = OracleGame(tree, o.ranges(), o.board_mask)
    for k, (r, s) in S0.items(): g.set_slab(k[0], k[1], r, s)
    if al is None: al = util.RowAligner(eng, g, tree)
    for k, (r, s) in S0.items(): al.write(k[0], k[1], r, s)
    eng.iterate_sampled([[c]]); g.iterate_sampled([[c]])
    bad = []
    for k in slabs:
        gr, gs = al.read(*k); orr, os_ = g.get_slab(*k)
        d = np.abs(gr - orr).max(); sc = max(np.abs(orr).max(), 1e-6)
        if d > 1e-4 * sc: bad.append((k, float(d), float(sc)))
    print('card', c, 'board', g.board_id_of([c])[0], 'bad', len(bad), bad[:3])
