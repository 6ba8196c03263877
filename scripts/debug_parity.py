import sys
import numpy as np
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from oracle import OracleGame
from tests import util

def run(name, o, card_abs=None, keys=None, iters=3):
    n, tree = rb.build_game_tree(o)
    r = o.ranges()
    eng = rb.Engine(tree, r, o.board_mask, card_abs or [], flags=rb.RS_FLAG_NO_GRAPH)
    orc = OracleGame(tree, r, o.board_mask, keys=keys)
    st = eng.stats()
    nb = [st.n_boards[k] for k in range(st.n_rounds)]
    print('==', name, 'rounds', st.n_rounds, 'boards', nb)
    for it in range(iters):
        eng.iterate(1); orc.iterate(1)
        rows = []
        scales = {}
        for an, b in util.all_slabs(tree, nb):
            gr, gs = eng.read_infoset(an, b)
            orr, os_ = orc.get_slab(an, b)
            for g, oo, nm in ((gr, orr, 'R'), (gs, os_, 'S')):
                if oo.size == 0: continue
                scales[(an, nm)] = max(scales.get((an, nm), 0), np.abs(oo).max())
                rows.append((an, b, nm, np.abs(g - oo).max(), np.abs(oo).max(), np.abs(g).max()))
        rows = [(d / max(scales[(an, nm)], 1e-30), an, b, nm, d, om, gm, scales[(an, nm)]) for an, b, nm, d, om, gm in rows]
        rows.sort(reverse=True)
        print(' iter', it, 'worst:')
        for x in rows[:6]:
            node = [i for i in range(tree.n_nodes) if tree.type[i] == 0 and tree.an_index[i] == x[1]][0]
            print('   rel %.3e an %d b %d %s absdiff %.3e |o| %.3e |g| %.3e scale %.3e  node %d P%d round %d' % (x + (node, tree.player[node], tree.round_idx[node])))
    return tree

o = util.small_options("4d5dAs3c", [util.RANGE_A, util.RANGE_B], [[0.5, 1.0]] * 2, [[3.0]] * 2)
run('turn_river', o)
o2 = util.small_options("4d5dAs", ["AA,KK,AKs,76s,54s", "QQ,JJ,AQs,65s,32s"], [[1.0]] * 3, [[3.0]] * 3, pot=40, stacks=(60, 60))
t = run('flop', o2, iters=2)
print("\n".join(t.dump()))
r = o.ranges()
k0 = util.bucket_keys_for(None, r, 1, 7, seed=3)
k1 = util.bucket_keys_for(None, r, 48, 11, seed=4)
abs_ = [rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k0), rb.CardAbstraction(rb.RS_ABS_BUCKET_TABLE, bucket_table=k1)]
run('bucketed', o, abs_, [k0, k1])
