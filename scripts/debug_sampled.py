import sys
import numpy as np
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from oracle import OracleGame
from tests import util
from tests.test_gpu_parity import _sample_paths
o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
n, tree = rb.build_game_tree(o)
eng = rb.Engine(tree, o.ranges(), o.board_mask, flags=rb.RS_FLAG_NO_GRAPH)
orc = OracleGame(tree, o.ranges(), o.board_mask)
al = util.RowAligner(eng, orc, tree)
rng = np.random.RandomState(101)
st = eng.stats(); nb = [st.n_boards[k] for k in range(st.n_rounds)]
for it in range(6):
    paths = _sample_paths(rng, o.board_mask, 1, 1)
    ids = [orc.board_id_of(p) for p in paths]
    if it > 0:
        util.copy_oracle_to_engine(eng, orc, tree, aligner=al)
    eng.iterate_sampled(paths); orc.iterate_sampled(paths)
    bad = []
    for an, b in util.all_slabs(tree, nb):
        gr, gs = al.read(an, b); orr, os_ = orc.get_slab(an, b)
        d = max(np.abs(gr - orr).max(), np.abs(gs - os_).max())
        if d > 1e-4 * max(np.abs(orr).max(), 1e-9) + 1e-9:
            bad.append((an, b, float(np.abs(gr - orr).max()), float(np.abs(orr).max()), float(np.abs(gr).max())))
    print('iter', it, 'paths', paths, 'ids', ids, 'bad slabs', len(bad), bad[:6])
    k_of = lambda an: int(tree.round_idx[util.node_of(tree, an)])
    print('   bad by round:', {k: sum(1 for x in bad if k_of(x[0]) == k) for k in (0, 1)}, 'bad boards:', sorted(set(x[1] for x in bad if k_of(x[0]) == 1))[:10])
