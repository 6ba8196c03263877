"""A/B of the chain-round task split (RS_FLAG_NO_CHAIN_SPLIT) on the bench workloads: ms per iteration, device time."""
import sys
sys.path.insert(0, '.')
import rustsolver_b200 as rb
from rustsolver_b200 import configs
for name in sys.argv[1:] or ["config1", "config2", "config3"]:
    w = getattr(configs, name)()
    n, tree = rb.build_game_tree(w.options)
    for label, fl in (("split", 0), ("one task per node", rb.RS_FLAG_NO_CHAIN_SPLIT)):
        eng = rb.Engine(tree, configs.workload_ranges(w), w.options.board_mask, w.card_abs, board_masks=w.board_masks, flags=fl)
        eng.iterate(20)
        ms0 = eng.stats().device_ms
        eng.iterate(100)
        ms = (eng.stats().device_ms - ms0) / 100
        print(f"{name:8s} {label:18s} {ms:8.4f} ms/iter  {1000 / ms:9.1f} iter/s (back to back, no L2 flush)")
        eng.close()
