"""Per-instance trace of the root street for one traversal (needs -DRS_TASK_TIMING): start/end of each (kind, round 0) task."""
import ctypes as C, sys
sys.path.insert(0, '.')
import numpy as np
import rustsolver_b200 as rb
from rustsolver_b200 import configs
name = sys.argv[1] if len(sys.argv) > 1 else 'config1'
w = getattr(configs, name)()
n, tree = rb.build_game_tree(w.options)
eng = rb.Engine(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, board_masks=w.board_masks, flags=rb.RS_FLAG_NO_GRAPH)
eng.iterate(5)
buf = np.zeros(96, dtype=np.uint64)
p64 = buf.ctypes.data_as(C.POINTER(C.c_uint64))
eng._lib.rs_debug_task_timing(eng._h, p64, 1)
eng.iterate(1)
eng._lib.rs_debug_task_timing(eng._h, p64, 1)
kinds = ['DOWN', 'UP_OPP', 'UP_TRAV', 'GATHER', 'ROOT_SD', 'CH_DOWN', 'CH_UP', 'TRAV_TERMS']
st = eng.stats()
print(name, 'ms/iter', st.device_ms / st.iterations, 'n action nodes', tree.n_actions)
for k in range(8):
    c, wait, body, whole = [int(x) for x in buf[4 * k:4 * k + 4]]
    if c: print(f'{kinds[k]:8s} n {c:6d}  body {body/c:9.0f} cyc = {body/c/1.9e3:6.2f} us')
rows = []
for k in range(8):
    for r in range(3):
        a, b = int(buf[32 + (k * 3 + r) * 2]), int(buf[32 + (k * 3 + r) * 2 + 1])
        if b: rows.append((a, b, kinds[k], r))
t0 = min(r[0] for r in rows)
for a, b, kd, r in sorted(rows): print(f'{kd:8s} round {r}: first start {(a - t0) / 1e3:8.1f}   last end {(b - t0) / 1e3:8.1f} us')

n = int(buf[80])
if n:
    print(f'TRAV phases: loads+scan+terms {int(buf[81])/n:8.0f} cyc, table update+store {int(buf[82])/n:8.0f} cyc')
m = int(buf[84])
if m:
    print(f'scan_reach inside trav_terms: {int(buf[85])/m:8.0f} cyc (n={m})')
