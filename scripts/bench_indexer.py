"""Device vs host hand indexer on the card-table workload of config 4 (2 352 river boards x 1 176 hands x 2 players)."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import rustsolver_b200 as rb
rng = np.random.default_rng(7)
n = 2352 * 1176 * 2
hands = np.argsort(rng.random((n // 8, 52)), axis=1)[:, :7].astype(np.uint8)
hands = np.tile(hands, (8, 1))
ix = rb.HandIndexer([2, 5])
ix.index_many_gpu(hands[:1000])  # context + tables warm-up
t0 = time.perf_counter(); dev, ms = ix.index_many_gpu(hands, return_ms=True); t_dev = time.perf_counter() - t0
t0 = time.perf_counter(); host = ix.index_many(hands); t_host = time.perf_counter() - t0
assert np.array_equal(dev, host)
print(json.dumps({"hands": int(len(hands)), "kernel_ms": ms, "kernel_hands_per_s": len(hands) / (ms * 1e-3),
                  "kernel_GBps_in_out": len(hands) * 15 / (ms * 1e-3) / 1e9,
                  "device_call_s_incl_copies": t_dev, "host_1_thread_s": t_host, "speedup_call": t_host / t_dev}))
