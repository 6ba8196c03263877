"""CPU: the oracle's restatement of emd_1d / l2_dist / Kmeans::predict against the reference's own known answers."""
import numpy as np

import oracle
from tests import abstraction_kats as K


def test_emd_1d_reference_known_answers():
    """emd.rs:122-180: identical histograms -> 0; 66 vs JT and 72 vs AA within ERROR = 0.01 of the exact EMD."""
    for p, q, want, exact in K.KATS:
        got = oracle.emd_1d(p, q)
        if exact:
            assert got == want
        else:
            assert abs(got - want) < K.ERROR, (got, want)
    assert oracle.emd_1d(K.H_66, K.H_66) == 0.0  # identical, normalised sum exactly 1: exactly 0
    # the one-ulp residue of test_same, step by step in numpy float32 (see tests/abstraction_kats.py)
    p = np.array(K.H_SAME, dtype=np.float32)
    s = np.float32(0)
    for v in p:
        s = np.float32(s + v)
    w = np.float32(0)
    for v in p:
        w = np.float32(w + np.float32(v / s))
    factor = min(max(np.float32(np.float32(4.45) * w) - np.float32(1.5), 1.0), 4.0)
    u = int(np.round(np.float32(len(p)) / np.float32(factor)))
    assert oracle.emd_1d(p, p) == float(np.float32(np.float32(1) - w) * np.float32(u))


def test_emd_1d_edge_cases():
    z = np.zeros(30, dtype=np.float32)
    assert oracle.emd_1d(z, K.H_66) == 0.0 and oracle.emd_1d(K.H_66, z) == 0.0  # emd.rs:60-62: an empty histogram
    one_hot = lambda i: np.eye(30, dtype=np.float32)[i]
    # all mass moves i -> j: the cross-bin search finds it at distance |i - j| while it is within reach (u = 30 here)
    assert oracle.emd_1d(one_hot(3), one_hot(10)) == 7.0
    assert oracle.emd_1d(one_hot(0), one_hot(29)) == 29.0
    # unnormalised input is normalised first (emd.rs:57-66)
    a = np.array(K.H_66, dtype=np.float32)
    assert abs(oracle.emd_1d(4 * a, K.H_JT) - oracle.emd_1d(a, K.H_JT)) < 1e-5


def test_l2_dist():
    a, b = np.array(K.H_66, dtype=np.float32), np.array(K.H_JT, dtype=np.float32)
    assert abs(oracle.l2_dist(a, b) - float(np.sqrt(((a - b) ** 2).sum()))) < 1e-6
    assert oracle.l2_dist(a, a) == 0.0


def test_kmeans_predict_is_the_first_nearest_centre():
    rng = np.random.default_rng(3)
    x = K.random_histograms(rng, 200, 30)
    c = K.random_histograms(rng, 17, 30)
    c[5] = c[11]  # a duplicated centre: the strict < of kmeans.rs:199 keeps the first
    for kind in (0, 1):
        cl, md, inertia = oracle.kmeans_predict(x, c, kind)
        f = oracle.emd_1d if kind == 0 else oracle.l2_dist
        for i in range(0, 200, 13):
            d = np.array([f(x[i], c[k]) for k in range(17)], dtype=np.float32)
            assert cl[i] == int(np.argmin(d)) and md[i] == d.min()
        assert not (cl == 11).any()
        assert abs(inertia - float(md.astype(np.float64).sum())) < 1e-9


def test_update_min_dists():
    rng = np.random.default_rng(4)
    x = K.random_histograms(rng, 50, 30)
    c = K.random_histograms(rng, 1, 30)[0]
    md0 = rng.random(50).astype(np.float32) * 20
    md = oracle.update_min_dists(x, c, md0, 0)
    want = np.minimum(md0, np.array([np.float32(oracle.emd_1d(x[i], c)) ** 2 for i in range(50)], dtype=np.float32))
    assert np.array_equal(md, want)


def _lloyd_reference(x, c, kind, rounds):
    """Plain Lloyd rounds with the reference's centre update (f32 sums in point order, kmeans.rs:522-543)."""
    c = c.copy()
    for _ in range(rounds):
        cl, _, _ = oracle.kmeans_predict(x, c, kind)
        mass = np.zeros_like(c)
        count = np.zeros(len(c), dtype=np.float32)
        for j in range(len(x)):
            count[cl[j]] += np.float32(1.0)
            mass[cl[j]] += x[j]
        for j in range(len(c)):
            pos = mass[j] > 0
            mass[j][pos] = mass[j][pos] / count[j]
        c = mass
    return cl, c


def test_fit_regular_first_round_is_a_full_assignment():
    """Round one: every upper bound is f32::MAX, so every point scans every centre -> Kmeans::predict, then the means."""
    rng = np.random.default_rng(21)
    x = K.random_histograms(rng, 300, 30)
    c0 = K.random_histograms(rng, 12, 30)
    for kind in (0, 1):
        cl, c1, inertia = oracle.kmeans_fit_regular(x, c0, kind, rounds=1)
        want_cl, want_c = _lloyd_reference(x, c0, kind, 1)
        assert np.array_equal(cl, want_cl)
        assert np.array_equal(c1, want_c)
        assert np.isfinite(inertia)


def test_fit_regular_with_a_metric_equals_plain_lloyd():
    """With l2_dist (a metric) the Hamerly bounds of kmeans.rs:285-334 only skip work: ten rounds give the assignments
    and centres of ten plain Lloyd rounds."""
    rng = np.random.default_rng(22)
    x = K.random_histograms(rng, 250, 16)
    c0 = x[rng.choice(len(x), 9, replace=False)].copy()
    cl, c, inertia = oracle.kmeans_fit_regular(x, c0, 1, rounds=10)
    want_cl, want_c = _lloyd_reference(x, c0, 1, 10)
    assert np.array_equal(cl, want_cl)
    assert np.allclose(c, want_c, rtol=0, atol=0)
    assert inertia >= 0


def test_kmeans_seeding_restatements():
    """Kmeans::init_pp / init_random (kmeans.rs:60-166) with the stated splitmix64 stream: reproducible, distinct centres,
    and k-means++ really prefers far points (three tight, far apart blobs: one centre lands in each)."""
    import oracle
    rng = np.random.default_rng(3)
    blobs = np.concatenate([np.abs(rng.normal(c, 0.01, size=(200, 8))) for c in (0.1, 1.0, 3.0)]).astype(np.float32)
    for seed in range(1, 6):
        ch = oracle.kmeans_init_pp(blobs, 3, 1, seed)
        assert np.array_equal(ch, oracle.kmeans_init_pp(blobs, 3, 1, seed))
        assert sorted(int(c) // 200 for c in ch) == [0, 1, 2], ch
    x = rng.random((500, 10)).astype(np.float32)
    ch = oracle.kmeans_init_random(x, 6, 9, 0, 4)
    assert len(set(ch.tolist())) == 6 and np.array_equal(ch, oracle.kmeans_init_random(x, 6, 9, 0, 4))
    # the winner is the most spread out of the restarts: its mean pairwise distance is the maximum over single-restart runs
    def spread(idx):
        c = x[idx]
        return np.mean([oracle.emd_1d(c[i], c[j]) for i in range(6) for j in range(6) if i != j])
    assert spread(ch) >= spread(oracle.kmeans_init_random(x, 6, 1, 0, 4)) - 1e-6


def test_fit_growbatch_is_one_mini_batch_step():
    """Kmeans::fit_growbatch (kmeans.rs:336-494) ends its loop with `break`: one pass over the first `batch` shuffled points.
    The restatement's assignment equals the nearest centre (first on ties) and the new centres are the batch means."""
    rng = np.random.default_rng(5)
    x = K.random_histograms(rng, 800, 20)
    c0 = x[rng.choice(len(x), 9, replace=False)].copy()
    for kind in (0, 1):
        idx, cl, c, min_change, inertia = oracle.kmeans_fit_growbatch(x, c0, 300, kind, seed=11)
        assert len(set(idx.tolist())) == 300 and idx.max() < len(x)  # a prefix of a permutation
        idx2, *_ = oracle.kmeans_fit_growbatch(x, c0, 800, kind, seed=11)
        assert sorted(idx2.tolist()) == list(range(len(x))) and np.array_equal(idx2[:300], idx)
        want, _, _ = oracle.kmeans_predict(x[idx], c0, kind)
        assert np.array_equal(cl, want)
        for j in range(len(c0)):
            members = x[idx][cl == j]
            if len(members):
                mean = members.astype(np.float64).mean(axis=0)
                assert np.allclose(c[j], mean, rtol=1e-5, atol=1e-7)
        assert np.isfinite(inertia) and inertia > 0 and min_change > 0


def test_generate_histograms_restatement():
    """generate_histograms (gen_abstraction/main.rs:79-159) with the EHS computed exactly: every row is a distribution over
    the bins; on the river nothing is left to draw, so a row is one spike at the bin of the hand's equity; a royal flush
    sits in the top bin, and the equities follow the evaluator's order."""
    import rustsolver_b200 as rb
    ix = rb.HandIndexer([2, 5])
    first, count, bins = 1000, 24, 30
    cards = np.zeros((count, 7), dtype=np.uint8)
    for i in range(count):
        cards[i] = ix.get_hand(1, first + i)
    h = oracle.generate_histograms(cards, 7, first, samples=5, bins=bins, seed=3)
    assert h.shape == (count, bins) and np.allclose(h.sum(axis=1), 1.0)
    assert ((h == 1.0).sum(axis=1) == 1).all()
    royal = np.array([[4 * 12, 4 * 11, 4 * 10, 4 * 9, 4 * 8, 1, 6]], dtype=np.uint8)  # As Ks | Qs Js Ts 2h 3c
    assert oracle.generate_histograms(royal, 7, 0, 3, bins, 1)[0, bins - 1] == 1.0
    # flop hands: 2 cards drawn per sample, rows are distributions, the stream depends on the hand index and the seed only
    ixf = rb.HandIndexer([2, 3])
    cf = np.zeros((8, 7), dtype=np.uint8)
    for i in range(8):
        cf[i, :5] = ixf.get_hand(1, 500 + i)
    a = oracle.generate_histograms(cf, 5, 500, samples=40, bins=10, seed=9)
    b = oracle.generate_histograms(cf[3:], 5, 503, samples=40, bins=10, seed=9)
    assert np.allclose(a.sum(axis=1), 1.0) and (a > 0).sum(axis=1).min() >= 2
    assert np.array_equal(a[3:], b)
    assert not np.array_equal(a, oracle.generate_histograms(cf, 5, 500, samples=40, bins=10, seed=10))
