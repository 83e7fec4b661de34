"""Tags out: phq_decode_batch_raw_tags writes, for every read, the BAM auxiliary bytes the reference's
Read::flush + Auxiliary::encode produce (read.h:187-237, auxiliary.cpp:320-361).

Checked against (a) the tags of the reference's own golden output test/BDGGG/valid/annotated.out (fixture
tests/golden/bdggg_expected.json: RG BC QT XB OX BZ CB CR CY XC, their order, short index reads included) and
(b) oracle/_ref, the reference's own Read / Auxiliary classes, on synthetic chains with reverse complemented and
knitted tokens, ragged reads, several decoders per topic and corrected molecular barcodes.
Bar: tag names, order and strings identical; float tags within 1e-6 relative (the confidence tolerance), floored at
the 2^-53 quantum of the reference's own f64 `1.0 - confidence`."""
import copy

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from pheniqs_b200 import DecoderChain, compile_job, workload
from pheniqs_b200.decoder import parse_auxiliary

pytestmark = pytest.mark.gpu

LETTER = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)
ORDER = ("RG", "BC", "QT", "XB", "RX", "QX", "OX", "BZ", "XM", "CB", "CR", "CY", "XC")


def raw_segments(code, quality, offset, phred_offset=33):
    """The FASTQ bytes the feed decoded (fastq.h:55-78), as phq_raw_segment tuples."""
    return [(LETTER[c], (q.astype(np.int32) + phred_offset).astype(np.uint8), o, 0) for c, q, o in zip(code, quality, offset)]


def compare(got, expected, label):
    assert list(got) == [t for t in ORDER if t in expected], "%s: tags %s, expected %s" % (label, list(got), sorted(expected))
    for tag, value in expected.items():
        if tag in ("XB", "XM", "XC"):
            # 1 - confidence is formed in f64 (read.h:189), so it is quantised at 2^-53 (same floor as helpers.compare_pamld)
            assert abs(float(got[tag]) - float(value)) <= 1e-6 * float(value) + 8 * 2.0 ** -53, "%s: %s = %r, expected %r" % (label, tag, got[tag], value)
        else:
            assert got[tag] == value, "%s: %s = %r, expected %r" % (label, tag, got[tag], value)


def test_bdggg_golden_tags():
    batch, decoders, expected = helpers.bdggg()
    compiled = helpers.golden("bdggg_compiled.json")        # the reference's own compile output: read group IDs included
    chain = DecoderChain(compiled, device=0)
    aux, length, flags = chain.decode_raw_tags(raw_segments(batch.code, batch.quality, batch.offset), batch.n_reads, 33, batch.qcfail)
    assert aux.shape[1] == chain.tag_record_bytes()
    for r, e in enumerate(expected):
        got = parse_auxiliary(aux[r, :length[r]])
        assert list(got) == e["order"], e["name"]
        for tag in ("RG", "BC", "QT", "OX", "BZ", "CB", "CR", "CY"):
            assert got.get(tag) == e[tag], "%s %s" % (e["name"], tag)
        for tag in ("XB", "XC"):
            assert (None if tag not in got else "%g" % got[tag]) == e[tag], "%s %s" % (e["name"], tag)
        assert (589 if flags[r] else 77) == e["flag"], e["name"]
        assert not aux[r, length[r]:].any()
    chain.close()


def chains(rng):
    job = {"sample": helpers.random_job(rng, "mdd", (6, 9), 32, minimum_distance=3),
           "molecular": [{"algorithm": "naive", "transform": {"token": ["0::4"]}}],
           "cellular": [helpers.random_job(rng, "pamld", (8,), 12, reverse=True), helpers.random_job(rng, "pamld", (5, 7), 20)]}
    job["cellular"][0]["transform"]["token"] = ["0:3:11"]
    job["cellular"][1]["transform"]["token"] = ["1:6:11", "0:1:8"]
    spelled = copy.deepcopy(job)
    spelled["cellular"][1]["transform"]["token"] = ["1:-5:", "0:1:8"]         # from the end of an 11 nt segment, open ended
    yield "mdd sample, naive umi, two pamld cellular", job, spelled
    knit = helpers.random_job(rng, "pamld", (10, 10), 30, **{"high quality threshold": 20, "high quality distance threshold": 2})
    knit["transform"] = {"token": ["0:0:6", "0:8:12", "1:2:12"], "knit": ["~0:1", "2"]}
    job = {"sample": helpers.random_job(rng, "pamld", (8, 8), 40), "molecular": [helpers.random_job(rng, "pamld", (9,), 16, **{"corrected quality": 17}), {"algorithm": "naive", "transform": {"token": ["1:0:5"]}}],
           "cellular": [knit, helpers.random_job(rng, "mdd", (7,), 10, minimum_distance=3)]}
    job["molecular"][0]["transform"]["token"] = ["1:3:12"]
    job["cellular"][1]["transform"]["token"] = ["0:5:12"]
    yield "pamld everywhere, corrected molecular barcode, knit, mdd cellular", job, job


@pytest.mark.parametrize("short", [0.0, 0.25])
@pytest.mark.parametrize("sub_batch", [0, 333])
def test_tags_equal_the_reference(short, sub_batch, monkeypatch):
    if not O.ref_available():
        pytest.skip("oracle/_ref is not built")
    if sub_batch:
        monkeypatch.setenv("PHQ_SUB_BATCH_READS", str(sub_batch))
    rng = np.random.default_rng(31)
    for label, job, spelled in chains(rng):
        n = 3000
        code, quality, offset, _ = workload.synthesize(compile_job(job), [0], n, seed=13, short_fraction=short)
        compiled = compile_job(spelled)
        qcfail = (rng.random(n) < 0.1).astype(np.uint8)
        checker = O.RefOracle(copy.deepcopy(compiled), len(code))
        expected, expected_flags = checker.tags(O.ReadBatch(code, quality, offset, qcfail))
        chain = DecoderChain(compiled, device=0)
        aux, length, flags, results = chain.decode_raw_tags(raw_segments(code, quality, offset), n, 33, qcfail, want_results=True)
        assert np.array_equal(flags, expected_flags), label
        for r in range(n):
            compare(parse_auxiliary(aux[r, :length[r]]), expected[r], "%s, read %d" % (label, r))
        # the same block from the reference's decoded form of the segments (phq_decode_batch_bam_tags)
        if not sub_batch:
            decoded = DecoderChain(compiled, device=0)
            aux2, length2, flags2 = decoded.decode_raw_tags([(c, q, o, 0) for c, q, o in zip(code, quality, offset)], n, 0, qcfail, bam=True)
            assert np.array_equal(aux2, aux) and np.array_equal(length2, length) and np.array_equal(flags2, flags), label
            decoded.close()
        # the tag path leaves the accumulators as any other decode call does
        for k, info in enumerate(chain.info):
            u, _ = chain.accumulators(k)
            eu, _ = checker.accumulators(k)
            assert np.array_equal(u, eu), label
        chain.close()


def test_stride_is_validated():
    rng = np.random.default_rng(5)
    compiled = compile_job({"sample": helpers.random_job(rng, "pamld", (8,), 12)})
    chain = DecoderChain(compiled, device=0)
    code, quality, offset, _ = workload.synthesize(compiled, [0], 10, seed=1)
    with pytest.raises(Exception):
        chain.decode_raw_tags(raw_segments(code, quality, offset), 10, 33, stride=8)
    chain.close()
