"""Gantt summary of ONE traversal per (task kind, round): needs a -DRS_TASK_TIMING build."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import numpy as np
import rustsolver_b200 as rb
from rustsolver_b200 import configs
name = sys.argv[1] if len(sys.argv) > 1 else 'config2'
w = getattr(configs, name)()
n, tree = rb.build_game_tree(w.options)
eng = rb.Engine(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, board_masks=w.board_masks, flags=rb.RS_FLAG_NO_GRAPH)
eng.iterate(5)
buf = np.zeros(96, dtype=np.uint64)
p64 = buf.ctypes.data_as(C.POINTER(C.c_uint64))
eng._lib.rs_debug_task_timing(eng._h, p64, 1)
# one traversal only: use best response of player 0? simpler: one full iteration = two traversals overlap in the stats,
# so profile a single EVAL traversal pair is not the CFR kernel; instead run iterate(1) and report both traversals merged
eng.iterate(1)
eng._lib.rs_debug_task_timing(eng._h, p64, 1)
kinds = ['DOWN', 'UP_OPP', 'UP_TRAV', 'GATHER', 'ROOT_SD', 'CH_DOWN', 'CH_UP', 'TRAV_TERMS']
rows = []
for k in range(8):
    for r in range(3):
        a, b = int(buf[32 + (k * 3 + r) * 2]), int(buf[32 + (k * 3 + r) * 2 + 1])
        if b:
            rows.append((a, b, kinds[k], r))
t0 = min(r[0] for r in rows)
print(name, '(times in us from the first task start; one iteration = traversal of player 0 then player 1, merged)')
for a, b, kd, r in sorted(rows):
    print(f'{kd:8s} round {r}: first start {(a - t0) / 1e3:8.1f}   last end {(b - t0) / 1e3:8.1f}')
