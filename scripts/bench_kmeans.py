"""k-means assignment step (Kmeans::predict, gen_abstraction/kmeans.rs:173-211) with emd_1d: device vs the CPU port."""
import sys, time, json
sys.path.insert(0, '.')
import numpy as np
import oracle
import rustsolver_b200 as rb
from tests import abstraction_kats as K
rng = np.random.default_rng(5)
n, k, dim = 200_000, 500, 50  # a slice of the flop round (1 286 792 canonical hands), 500 clusters, 50 bins
x = K.random_histograms(rng, n, dim)
c = K.random_histograms(rng, k, dim)
rb.kmeans_assign(x[:1000], c[:4])
t0 = time.perf_counter(); cl, md, inertia, ms = rb.kmeans_assign(x, c, rb.RS_DIST_EMD_1D, return_ms=True); t_dev = time.perf_counter() - t0
m = 4000  # bounded CPU sample
t0 = time.perf_counter(); ocl, omd, _ = oracle.kmeans_predict(x[:m], c, 0); t_cpu = time.perf_counter() - t0
assert np.array_equal(cl[:m], ocl) and np.array_equal(md[:m], omd)
print(json.dumps({"points": n, "centres": k, "bins": dim, "kernel_ms": ms, "pairs_per_s_kernel": n * k / (ms * 1e-3),
                  "device_call_s": t_dev, "cpu_port_pairs_per_s": m * k / t_cpu, "cpu_threads": "all (OpenMP)",
                  "speedup_kernel_vs_cpu": (n * k / (ms * 1e-3)) / (m * k / t_cpu)}))
