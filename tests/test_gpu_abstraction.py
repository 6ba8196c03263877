"""GPU: emd_1d / l2_dist / k-means assignment kernels (abstraction_kernels.cu) through the C ABI against the
reference's known answers and, bit for bit, against the CPU restatement (same fp32 operations in the same order)."""
import numpy as np
import pytest

import oracle
import rustsolver_b200 as rb
from tests import abstraction_kats as K

pytestmark = pytest.mark.gpu


def test_emd_1d_reference_known_answers_on_the_device():
    for p, q, want, exact in K.KATS:
        got = float(rb.histogram_distances(np.array([p], np.float32), np.array([q], np.float32))[0])
        assert (got == want) if exact else (abs(got - want) < K.ERROR), (got, want)
        assert got == oracle.emd_1d(p, q)  # and identical to the CPU restatement


@pytest.mark.parametrize("dim", [8, 30, 50, 128])
def test_distances_are_bit_exact(dim):
    rng = np.random.default_rng(dim)
    p = K.random_histograms(rng, 4000, dim)
    q = K.random_histograms(rng, 4000, dim)
    p[7] = 0.0  # an empty histogram: distance 0 (emd.rs:60-62)
    q[9] = p[9]
    for kind, f in ((rb.RS_DIST_EMD_1D, oracle.emd_1d), (rb.RS_DIST_L2, oracle.l2_dist)):
        dev = rb.histogram_distances(p, q, kind)
        host = np.array([f(p[i], q[i]) for i in range(len(p))], dtype=np.float32)
        assert np.array_equal(dev, host), np.abs(dev - host).max()
    assert rb.histogram_distances(p, q, rb.RS_DIST_EMD_1D)[7] == 0.0


@pytest.mark.parametrize("n,k,dim", [(1, 1, 30), (63, 5, 30), (5000, 200, 50), (4097, 33, 8)])
def test_kmeans_assignment_matches_predict(n, k, dim):
    rng = np.random.default_rng(n + k)
    x = K.random_histograms(rng, n, dim)
    c = K.random_histograms(rng, k, dim)
    if k > 3:
        c[k - 1] = c[1]  # duplicated centre: the first one wins (strict <, kmeans.rs:199)
    for kind in (rb.RS_DIST_EMD_1D, rb.RS_DIST_L2):
        cl, md, inertia = rb.kmeans_assign(x, c, kind)
        ocl, omd, oin = oracle.kmeans_predict(x, c, kind)
        assert np.array_equal(cl, ocl)
        assert np.array_equal(md, omd)
        assert abs(inertia - oin) <= 1e-9 * max(abs(oin), 1.0)
    assert len(rb.kmeans_assign(x[:0], c)[0]) == 0  # empty data set


def test_update_min_dists_matches():
    rng = np.random.default_rng(11)
    x = K.random_histograms(rng, 3000, 30)
    c = K.random_histograms(rng, 1, 30)[0]
    md0 = (rng.random(3000) * 30).astype(np.float32)
    assert np.array_equal(rb.kmeans_update_min_dists(x, c, md0), oracle.update_min_dists(x, c, md0, 0))


@pytest.mark.parametrize("kind", [rb.RS_DIST_EMD_1D, rb.RS_DIST_L2])
def test_fit_regular_matches_the_cpu_restatement(kind):
    """Kmeans::fit_regular (kmeans.rs:497-599): ten rounds with the reference's bounds, device vs CPU, bit for bit."""
    rng = np.random.default_rng(31 + kind)
    x = K.random_histograms(rng, 3000, 30)
    c0 = x[rng.choice(len(x), 20, replace=False)].copy()
    cl, c, inertia = rb.kmeans_fit_regular(x, c0, kind, rounds=10)
    ocl, oc, oin = oracle.kmeans_fit_regular(x, c0, kind, rounds=10)
    assert np.array_equal(cl, ocl)
    assert np.array_equal(c, oc)
    assert inertia == oin
    assert len(np.unique(cl)) > 5


@pytest.mark.parametrize("kind", [rb.RS_DIST_EMD_1D, rb.RS_DIST_L2])
def test_kmeans_seeding_matches_the_cpu_restatement(kind):
    """Kmeans::init_pp / init_random (kmeans.rs:60-166) on the device against the CPU restatement with the same stated
    random stream: the same points are chosen (the distances under the draws are bit-identical)."""
    import oracle
    rng = np.random.default_rng(5)
    x = rng.random((3000, 20)).astype(np.float32)
    x /= x.sum(axis=1, keepdims=True)
    for seed in (1, 2):
        chosen, centers = rb.kmeans_init_pp(x, 25, kind, seed)
        assert np.array_equal(chosen, oracle.kmeans_init_pp(x, 25, kind, seed))
        assert np.array_equal(centers, x[chosen]) and len(set(chosen.tolist())) > 20
        chosen, centers = rb.kmeans_init_random(x, 12, 7, kind, seed)
        assert np.array_equal(chosen, oracle.kmeans_init_random(x, 12, 7, kind, seed))
        assert np.array_equal(centers, x[chosen]) and len(set(chosen.tolist())) == 12


@pytest.mark.parametrize("kind", [rb.RS_DIST_EMD_1D, rb.RS_DIST_L2])
def test_fit_growbatch_matches_the_cpu_restatement(kind):
    """Kmeans::fit_growbatch (kmeans.rs:336-494; one pass because its loop ends with `break`): shuffle, init_s,
    assignment_with_bounds of the first batch, mean update -- device vs CPU restatement, bit for bit."""
    rng = np.random.default_rng(77 + kind)
    x = K.random_histograms(rng, 5000, 30)
    c0 = x[rng.choice(len(x), 24, replace=False)].copy()
    for batch, seed in ((1000, 3), (5000, 4)):
        idx, cl, c, mc, inertia = rb.kmeans_fit_growbatch(x, c0, batch, kind, seed)
        oidx, ocl, oc, omc, oin = oracle.kmeans_fit_growbatch(x, c0, batch, kind, seed)
        assert np.array_equal(idx, oidx)
        assert np.array_equal(cl, ocl)
        assert np.array_equal(c, oc)
        assert mc == omc and inertia == oin
        assert len(np.unique(cl)) > 5


@pytest.mark.parametrize("round_,first,count,samples,bins", [(0, 0, 169, 60, 30), (1, 12345, 40, 50, 50), (2, 999999, 32, 40, 30), (3, 5000000, 48, 3, 20)])
def test_generate_histograms_matches_the_cpu_restatement(round_, first, count, samples, bins):
    """generate_histograms (gen_abstraction/main.rs:79-159) on the device, EHS computed exactly: bit for bit the CPU
    restatement's histograms on the hands the product un-indexed (the indexer is pinned separately, tests/test_poker.py)."""
    h, cards = rb.generate_histograms(round_, first, count, samples, bins, seed=21)
    n_known = [2, 5, 6, 7][round_]
    want = oracle.generate_histograms(cards, n_known, first, samples, bins, seed=21)
    assert np.array_equal(h, want), np.abs(h - want).max()
    assert np.allclose(h.sum(axis=1), 1.0)
    ix = rb.HandIndexer([2] if round_ == 0 else [2, [0, 3, 4, 5][round_]])
    assert list(cards[3, :n_known]) == list(ix.get_hand(0 if round_ == 0 else 1, first + 3))
