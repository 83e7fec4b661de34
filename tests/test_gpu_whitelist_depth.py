"""Config 5 at depth: the pruned whitelist scan (pamld_whitelist_kernel) against the oracle at the FULL table size,
737,280 x 16 nt, on read sets built to stress what the pruning approximates.

The whitelist kernel drops (read, barcode) pairs whose prior adjusted product is provably below
min(best / 2, 2^-21 (noise term + rest) / N): assignments are exact, and sigma_p loses at most 2^-21 of its non-winner
part, i.e. confidence and 1 - confidence move by < 4.8e-7 relative in the worst case. These tests
measure how much of that budget the CUDA code actually uses where it is most exposed:

  synthetic     20,000 reads of the bench distribution (the oracle's 32 cores need about half a minute)
  all Q12/Q23   every base at one low quality: no position is confidently called, the bound is at its loosest and
                sigma_p is spread over thousands of barcodes
  twins         reads exactly between two whitelist entries at Hamming distance 2 (one mismatch from each), the
                differing positions called at equal and at unequal qualities: near ties that must not be pruned
  no floor      a decoder with `noise: 0` (the absolute floor of the threshold vanishes) and unequal priors
  qcfail in     incoming chastity flags on a tenth of the reads

Every set is compared read by read (index, distance, qcfail exact; ln confidence and the error probability within
1e-6) and through the accumulators (u64 exact, f64 within 1e-9; for the all-low-quality sets within the provable
2^-21 x (1 - confidence threshold), see there); the maxima are printed so the run records the slack."""
import os

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from pheniqs_b200 import DecoderChain, compile_job, workload

pytestmark = pytest.mark.gpu

BAM = np.array([1, 2, 4, 8], dtype=np.uint8)


@pytest.fixture(scope="module")
def table():
    return workload.whitelist()                 # [737280, 16] 2-bit codes, barcode index order


def reads_from(barcode_codes, quality, rng, substitute=True):
    """[n, 28] BAM codes and qualities: the 16 barcode bases called at `quality` (substituted with probability
    10^(-q/10) when asked to), then a 12 nt UMI."""
    n = barcode_codes.shape[0]
    base = barcode_codes.astype(np.int64)
    if substitute:
        error = rng.random(base.shape) < np.power(10.0, -quality[:, :16].astype(np.float64) / 10.0)
        base = np.where(error, (base + rng.integers(1, 4, size=base.shape)) % 4, base)
    code = np.concatenate([BAM[base], BAM[rng.integers(0, 4, size=(n, 12))]], axis=1)
    return code, quality


def run(compiled, code, quality, qcfail=None, label="", accumulator_gate=1e-9):
    n = code.shape[0]
    batch = O.ReadBatch.from_fixed([code, np.zeros((n, 0), dtype=np.uint8)], [quality, np.zeros((n, 0), dtype=np.uint8)], qcfail)
    chain = DecoderChain(compiled, device=0)
    k = next(i for i, info in enumerate(chain.info) if info.algorithm == 0)
    assert "pamld_whitelist_kernel" in chain.kernel_description(k)
    results, flags = chain.decode(chain.pack(batch.code, batch.quality, batch.offset), n, qcfail)
    checker = O.best_oracle(compiled, 2)
    expected = checker.decode(batch, threads=os.cpu_count() or 1)
    got = results[k]
    helpers.compare_pamld(got, expected.index[:, k], expected.distance[:, k], expected.confidence[:, k], label)
    assert np.array_equal(flags, expected.qcfail), label
    u, f = chain.accumulators(k)
    eu, ef = checker.accumulators(k)
    assert np.array_equal(u, eu), label + " integer accumulators"
    relative = np.abs(f - ef) / np.maximum(np.abs(ef), 1e-300)
    relative[ef == 0] = np.abs(f[ef == 0])
    assert relative.max() <= accumulator_gate, label + " confidence accumulators: %g" % relative.max()
    ok = expected.confidence[:, k] > 0
    ln_error = np.abs(np.log(got["confidence"][ok]) - np.log(expected.confidence[ok, k])).max() if ok.any() else 0.0
    e_ref = 1.0 - expected.confidence[ok, k]
    visible = e_ref > 1e-9
    e_error = (np.abs((1.0 - got["confidence"][ok]) - e_ref)[visible] / e_ref[visible]).max() if visible.any() else 0.0
    print("\n%s: %d reads, %d classified, %d on the exact tie path, max |d ln confidence| %.3g, max relative d(1 - confidence) %.3g, max relative accumulator error %.3g" % (
        label, n, int((got["index"] > 0).sum()), chain.statistics()["exact_path_reads"], ln_error, e_error, relative.max()))
    chain.close()
    return got


def test_full_table_at_depth(table):
    spec = workload.load("c5")
    compiled = compile_job(spec["job"])
    rng = np.random.default_rng(2024)

    # the bench distribution, 20,000 reads
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], 20000, seed=12)
    run(compiled, code[0].reshape(20000, 28), quality[0].reshape(20000, 28), label="synthetic")

    # every base at one low quality
    for q in (12, 23):
        n = 1500
        drawn = rng.integers(0, table.shape[0], size=n)
        code, quality = reads_from(table[drawn], np.full((n, 28), q, dtype=np.uint8), rng)
        # every read of this set has 1 - confidence of a few percent, so the pruned mass (at most 2^-21 of the non-winner
        # part of sigma_p per read) is bounded by 2^-21 x (1 - confidence threshold) = 2.4e-8 of an accumulated
        # confidence, not by the 1e-9 ordinary read sets meet with three orders of margin: the gate here is that bound
        got = run(compiled, code, quality, label="all Q%d" % q, accumulator_gate=2.0 ** -21 * 0.05)
        assert (got["index"] > 0).any()

    # twins: a read between two whitelist entries at distance 2
    words = (table.astype(np.int64) << (2 * np.arange(15, -1, -1, dtype=np.int64))[None, :]).sum(axis=1)
    pairs = []
    for i, j in ((0, 1), (5, 11), (14, 15), (3, 9)):
        masked = words & ~((3 << (2 * (15 - i))) | (3 << (2 * (15 - j))))
        order = np.argsort(masked, kind="stable")
        same = np.nonzero(masked[order][1:] == masked[order][:-1])[0]
        for s in same:
            a, b = int(order[s]), int(order[s + 1])
            if table[a, i] != table[b, i] and table[a, j] != table[b, j]:
                pairs.append((a, b, i, j))
    assert len(pairs) > 200
    pairs = pairs[:600]
    n = len(pairs) * 2
    middle = np.zeros((n, 16), dtype=np.uint8)
    quality = np.full((n, 28), 37, dtype=np.uint8)
    for r, (a, b, i, j) in enumerate(pairs):
        for variant in (0, 1):
            row = table[a].copy()
            row[j] = table[b, j]                        # position i from a, position j from b: one mismatch from each
            middle[2 * r + variant] = row
            if variant == 1:
                quality[2 * r + 1, i] = 23              # unequal qualities at the two deciding positions
                quality[2 * r + 1, j] = 12
    code, quality = reads_from(middle, quality, rng, substitute=False)
    got = run(compiled, code, quality, label="twins")
    expected_pairs = np.array([[a + 1, b + 1] for a, b, _, _ in pairs]).repeat(2, axis=0)
    # (a third whitelist entry can sit as close: about one read in a hundred; the oracle comparison above is the check)
    assert np.mean((got["index"] == expected_pairs[:, 0]) | (got["index"] == expected_pairs[:, 1])) > 0.9

    # incoming qcfail flags
    n = 1500
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], n, seed=13)
    qcfail = (rng.random(n) < 0.1).astype(np.uint8)
    run(compiled, code[0].reshape(n, 28), quality[0].reshape(n, 28), qcfail, label="qcfail in")


def test_full_table_without_noise_floor_and_unequal_priors(table):
    spec = workload.load("c5")
    rng = np.random.default_rng(7)
    job = spec["job"]
    job["cellular"][0]["noise"] = 0.0
    weights = rng.integers(1, 6, size=table.shape[0])
    for record, weight in zip(job["cellular"][0]["codec"].values(), weights):
        record["concentration"] = float(weight)
    compiled = compile_job(job)
    n = 3000
    drawn = rng.integers(0, table.shape[0], size=n)
    quality = np.array([37, 23, 12, 2], dtype=np.uint8)[rng.choice(4, size=(n, 28), p=[0.7, 0.12, 0.12, 0.06])]
    code, quality = reads_from(table[drawn], quality, rng)
    code[:300, :16] = BAM[rng.integers(0, 4, size=(300, 16))]        # reads that belong to no barcode: nothing but the bound stops the scan
    run(compiled, code, quality, label="noise 0, unequal priors")
