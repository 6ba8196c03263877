"""Per-task-kind cycle breakdown (needs a library built with -DRS_TASK_TIMING; see scripts/build_variant.sh)."""
import ctypes as C
import sys
sys.path.insert(0, '.')
import numpy as np
import rustsolver_b200 as rb
from rustsolver_b200 import configs
name = sys.argv[1] if len(sys.argv) > 1 else 'config2'
w = getattr(configs, name)()
n, tree = rb.build_game_tree(w.options)
eng = rb.Engine(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, board_masks=w.board_masks)
eng.iterate(5)
buf = np.zeros(32, dtype=np.uint64)
eng._lib.rs_debug_task_timing(eng._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
iters = 10
eng.iterate(iters)
eng._lib.rs_debug_task_timing(eng._h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 1)
st = eng.stats()
kinds = ['DOWN', 'UP_OPP', 'UP_TRAV', 'GATHER', 'ROOT_SD', 'CH_DOWN', 'CH_UP', 'TRAV_TERMS']
print(name, 'ms/iter', st.device_ms / st.iterations)
tot = 0
for k in range(8):
    c, wait, body, whole = [int(x) for x in buf[4 * k:4 * k + 4]]
    if c:
        print(f'{kinds[k]:8s} n/iter {c/iters:9.0f}  wait {wait/c:9.0f} cyc  body {body/c:9.0f} cyc  whole {whole/c:9.0f} cyc   share of CTA-time {whole}')
        tot += whole
for k in range(8):
    c, wait, body, whole = [int(x) for x in buf[4 * k:4 * k + 4]]
    if c:
        print(f'{kinds[k]:8s} {100*whole/tot:5.1f}% of CTA time ({100*wait/tot:5.1f}% waiting, {100*body/tot:5.1f}% body)')
