"""Writes tests/golden/workload_<name>.npz: everything bench.py's reference arm needs to rebuild a BASELINE workload
WITHOUT importing the product -- the Options fields (the tree itself is rebuilt by oracle/tree_oracle.py), the
board-filtered ranges in the product's hand order, the subgame boards of config 5 and the bucket keys of the abstracted
root round (cluster_arr looked up through the hand indexer, card_abstraction.rs:204-209).  Regenerate only when
rustsolver_b200/configs.py changes:  python scripts/make_workload_fixtures.py"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import rustsolver_b200 as rb  # noqa: E402
from rustsolver_b200 import configs  # noqa: E402

G = ROOT / "tests" / "golden"
for name in ("config1", "config2", "config3", "config4", "config5"):
    w = getattr(configs, name)()
    o = w.options
    ranges = configs.workload_ranges(w)
    aa = o.action_abstraction
    meta = dict(name=w.name, stack_sizes=list(o.stack_sizes), board_mask=int(o.board_mask), starting_pot=int(o.starting_pot),
                bet_sizes=aa.bet_sizes, raise_sizes=aa.raise_sizes, abstraction=[int(a.kind) for a in w.card_abs])
    arrays = {"range0": np.asarray(ranges[0], dtype=np.uint8), "range1": np.asarray(ranges[1], dtype=np.uint8)}
    if w.board_masks:
        arrays["board_masks"] = np.asarray(w.board_masks, dtype=np.uint64)
    if w.card_abs and w.card_abs[0].kind == rb.RS_ABS_CLUSTER_ARR:
        board = [c for c in range(52) if o.board_mask >> c & 1]
        ix = rb.HandIndexer([2, len(board)])
        for q in range(2):
            cards = np.zeros((len(ranges[q]), 2 + len(board)), dtype=np.uint8)
            cards[:, :2] = np.sort(np.asarray(ranges[q], dtype=np.uint8), axis=1)
            cards[:, 2:] = board
            arrays[f"keys{q}"] = w.card_abs[0].cluster_arr[ix.index_many(cards)].astype(np.uint32)[None, :]
    np.savez_compressed(G / f"workload_{name}.npz", meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8), **arrays)
    print(name, {k: v.shape for k, v in arrays.items()})
