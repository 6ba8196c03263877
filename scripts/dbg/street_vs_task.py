import sys, numpy as np
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from oracle import OracleGame
from tests import util
o = util.small_options("4d5dAs3c", ["random", "random"], [[0.5, 1.0]] * 2, [[3.0]] * 2)
n, tree = rb.build_game_tree(o)
r = o.ranges()
e1 = rb.Engine(tree, r, o.board_mask, flags=rb.RS_FLAG_STREET_KERNEL)
e2 = rb.Engine(tree, r, o.board_mask)
orc = OracleGame(tree, r, o.board_mask)
st = e1.stats()
nb = [st.n_boards[k] for k in range(st.n_rounds)]
al1 = util.RowAligner(e1, orc, tree); al2 = util.RowAligner(e2, orc, tree)
for it in range(2):
    if it > 0:
        util.copy_oracle_to_engine(e1, orc, tree, aligner=al1)
        util.copy_oracle_to_engine(e2, orc, tree, aligner=al2)
    e1.iterate(1); e2.iterate(1); orc.iterate(1)
    rows = []
    for an, b in util.all_slabs(tree, nb):
        g1 = al1.read(an, b); g2 = al2.read(an, b); oo = orc.get_slab(an, b)
        for w in range(2):
            sc = float(np.abs(oo[w]).max())
            d1 = float(np.abs(g1[w] - oo[w]).max()); d2 = float(np.abs(g2[w] - oo[w]).max())
            rows.append((d1, d2, sc, an, b, w))
    tab = max(x[2] for x in rows)
    rows.sort(key=lambda x: -x[0])
    print("iter", it, "table max", tab)
    for x in rows[:8]:
        d1, d2, sc, an, b, w = x
        node = util.node_of(tree, an)
        print(f"  street-vs-orc {d1:.3e} task-vs-orc {d2:.3e} slab scale {sc:.3e} an {an} board {b} arr {w} player {tree.player[node]} round {tree.round_idx[node]}")
    # where inside the worst slab
    d1, d2, sc, an, b, w = rows[0]
    g1 = al1.read(an, b)[w]; oo = orc.get_slab(an, b)[w]
    idx = np.argsort(-np.abs(g1 - oo).max(axis=1))[:6]
    print("   worst rows", idx, np.abs(g1 - oo).max(axis=1)[idx])
    print("   board mask", [c for c in range(52) if orc.board_mask(1, b) >> c & 1])
    q = int(tree.player[util.node_of(tree, an)])
    for i in idx[:3]:
        slot = int(np.nonzero(orc.rows(1, q, b) == i)[0][0])
        print("    row", i, "hand", r[q][slot], "street", g1[i], "oracle", oo[i])
