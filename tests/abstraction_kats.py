"""The reference's own known answers for emd_1d (/root/reference/src/gen_abstraction/emd.rs:122-180), copied as data:
three histogram pairs, the expected value and the tolerance its tests assert (ERROR = 0.01, emd.rs:120)."""
H_SAME = [0.007493939393939393, 0.019696969696969702, 0.04244242424242425, 0.04021212121212122, 0.0871090909090909,
          0.05862121212121213, 0.040224242424242426, 0.0962121212121212]
H_66 = [0.0, 0.0, 0.0005, 0.0065, 0.0025, 0.0005, 0.0065, 0.003, 0.0115, 0.0095, 0.0135, 0.023, 0.012, 0.038, 0.0705, 0.0625,
        0.0725, 0.082, 0.1005, 0.052, 0.036, 0.036, 0.047, 0.023, 0.025, 0.022, 0.0355, 0.035, 0.04, 0.1335]
H_JT = [0.0035, 0.008, 0.0085, 0.0205, 0.034, 0.032, 0.007, 0.043, 0.0875, 0.0075, 0.036, 0.0405, 0.0175, 0.017, 0.025, 0.036,
        0.009, 0.0095, 0.0145, 0.0245, 0.057, 0.056, 0.055, 0.035, 0.0395, 0.0215, 0.042, 0.042, 0.057, 0.114]
H_27 = [0.054, 0.151, 0.0345, 0.12, 0.014, 0.012, 0.0095, 0.007, 0.0135, 0.018, 0.0185, 0.0455, 0.0835, 0.014, 0.03, 0.05,
        0.057, 0.0395, 0.018, 0.012, 0.0185, 0.0175, 0.0105, 0.009, 0.01, 0.0135, 0.046, 0.0405, 0.009, 0.024]
H_AA = [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0015, 0.0015, 0.0035, 0.0, 0.0, 0.0085, 0.0215, 0.0065, 0.0085,
        0.0015, 0.025, 0.0385, 0.0125, 0.026, 0.097, 0.21, 0.1455, 0.111, 0.082, 0.1995]
ERROR = 0.01
# (p, q, expected, exact?)  The reference's test_same (a #[bench], only run by `cargo bench`) asserts equality with 0.0,
# but its own arithmetic cannot give that: in IEEE fp32 the eight normalised bins of H_SAME sum to 1 - 2^-24, so
# emd_1d returns (1 - w) * u = 1.8e-7 (numpy float32 reproduces this step by step).  It is therefore checked like the
# other two, |emd - expected| < ERROR; identical histograms whose normalised sum IS exactly 1 give exactly 0.0
# (H_66 vs H_66, tests/test_abstraction.py).
KATS = [(H_SAME, H_SAME, 0.0, False), (H_66, H_JT, 2.709499043500001, False), (H_27, H_AA, 14.220495694500006, False)]


def random_histograms(rng, n, dim, sparsity=0.3):
    """Normalised histograms like generate_histograms produces (gen_abstraction/main.rs:79-159): counts / samples,
    with a share of empty bins."""
    import numpy as np
    counts = rng.integers(0, 40, size=(n, dim)).astype(np.float32)
    counts[rng.random((n, dim)) < sparsity] = 0.0
    s = counts.sum(axis=1, keepdims=True)
    s[s == 0] = 1.0
    return (counts / s).astype(np.float32)
