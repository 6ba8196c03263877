"""Where a task's time goes (needs a -DRS_TASK_TIMING build: scripts/build_variant.sh timing -DRS_TASK_TIMING).

Cycles of thread 0 of the team, per task kind and per phase inside the two heavy kinds."""
import ctypes as C, sys
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
it = 5
ms0 = eng.stats().device_ms
eng.iterate(it)
eng._lib.rs_debug_task_timing(eng._h, p64, 1)
st = eng.stats()
kinds = ['DOWN', 'UP_OPP', 'UP_TRAV', 'GATHER', 'ROOT_SD', 'CH_DOWN', 'CH_UP', '?']
print(name, 'ms/iter', (st.device_ms - ms0) / it)
cnt = {}
for k in range(8):
    c, wait, body, whole = [int(x) for x in buf[4 * k:4 * k + 4]]
    cnt[kinds[k]] = c
    if c: print(f'{kinds[k]:8s} n/iter {c/it:7.0f}  body {body/c:9.0f} cyc = {body/c/1.9e3:6.2f} us   total {body/it/1.9e6:8.2f} ms of team time per iter')
ph = [int(x) for x in buf[80:88]]
nt, nd = max(cnt['UP_TRAV'] + cnt['ROOT_SD'], 1), max(cnt['DOWN'], 1)
print(f'UP_TRAV per task: stage reach {ph[0]/nt:8.0f}  scan {ph[1]/nt:8.0f}  terms+update {ph[2]/max(cnt["UP_TRAV"],1):8.0f} cyc')
print(f'DOWN    per task: sigma+reach {ph[4]/nd:8.0f}  scans {ph[5]/nd:8.0f}  terminal terms {ph[6]/nd:8.0f} cyc')
