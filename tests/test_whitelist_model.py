"""The arithmetic pamld_whitelist_kernel relies on, restated in numpy and checked as properties on random
reads and barcodes. This is a model of the pruning rule, not of the CUDA code (that is what tests/test_gpu_parity.py
checks against the oracle): it pins the inequalities the rule is built on, so a change of the rule has something to fail.

  * bound[c] >= p_b for every barcode with c counted mismatches, whatever prefix of the weakest positions is left out;
  * with limit = max{c : bound[c] >= threshold}, a pruned barcode is below half the maximum (never the winner, never a
    tie) and the pruned mass is within `tolerance` of (noise term + rest)."""
import numpy as np

B = 10.0 ** -0.1
TOLERANCE = 2.0 ** -21


def ratio(q):
    """mismatch ratio of one position: B^(q - true positive quality(q)), phred.cpp:24-72"""
    if q == 0:
        return 1.0
    true_positive_quality = -10.0 * np.log10(1.0 - B ** q)
    return B ** (q - true_positive_quality)


def test_bound_dominates_and_pruning_stays_within_tolerance():
    rng = np.random.default_rng(12)
    L, N = 16, 20000
    barcodes = rng.integers(0, 4, size=(N, L))
    prior = np.full(N, 0.99 / N)
    noise_term = 0.01 * 4.0 ** -L
    qualities = np.array([37, 23, 12, 2])
    for trial in range(60):
        read = barcodes[rng.integers(N)].copy() if trial % 4 else rng.integers(0, 4, size=L)      # every fourth read is noise
        q = rng.choice(qualities, size=L, p=[0.85, 0.08, 0.05, 0.02])
        ambiguous = (q == 2) & (rng.random(L) < 0.5)
        error = rng.random(L) < np.power(10.0, -q / 10.0)
        read[error] = (read[error] + rng.integers(1, 4, size=int(error.sum()))) % 4
        w = np.array([1.0 if a else ratio(x) for x, a in zip(q, ambiguous)])
        mismatch = (barcodes != read[None, :]) | ambiguous[None, :]
        p = prior * np.prod(np.where(mismatch, w[None, :], 1.0), axis=1)       # relative to P0, as the kernels work

        candidates = np.flatnonzero((w < 1.0) & ~ambiguous)
        order = candidates[np.argsort(-w[candidates], kind="stable")]         # weakest (largest ratio) first
        loose = np.prod(w[w > 1.0])
        best, rest = p.max(), p.sum() - p.max()
        threshold = min(0.5 * best, TOLERANCE * (noise_term + rest) / N)
        for skipped in range(len(order) + 1):
            counted = order[skipped:]
            bound = prior.max() * loose * np.concatenate([[1.0], np.cumprod(w[counted])]) * (1.0 + 2.0 ** -20)
            c = mismatch[:, counted].sum(axis=1)
            assert np.all(p <= bound[c]), "a barcode exceeds the bound of its mismatch count"
            limit = max(k for k in range(len(counted) + 1) if bound[k] >= threshold)
            pruned = c > limit
            assert not pruned[np.argmax(p)]
            assert np.all(p[pruned] < 0.5 * best)
            assert p[pruned].sum() <= TOLERANCE * (noise_term + rest)
