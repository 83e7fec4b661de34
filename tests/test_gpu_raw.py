"""Feed bytes in: phq_decode_batch_raw packs on the device what phq_pack packs on the host.
Parity bar: every tile bit — hence every result, flag and accumulator — identical to the host packed path, including
the state short tokens leave behind across sub-batches and across calls.

Anchoring: the first three tests compare the device packer with the host packer (product against product); the host
packer is pinned on the oracle's Rule::apply on the CPU (tests/test_abi.py) and the raw path on oracle/_ref in
tests/test_gpu_tags.py. test_bam_segments_equal_the_oracle compares the BAM-code form with the oracle directly."""
import copy
import os

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from pheniqs_b200 import DecoderChain, compile_job, workload

pytestmark = pytest.mark.gpu

LETTER = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)


def to_fastq_bytes(code, quality, phred_offset=33):
    """What the FASTQ records held before the feed decoded them (fastq.h:55-78)."""
    return [LETTER[c] for c in code], [(q.astype(np.int32) + phred_offset).astype(np.uint8) for q in quality]


def jobs(rng):
    """(label, job to synthesize reads from, job to decode with): the second may spell tokens differently."""
    job = {"sample": helpers.random_job(rng, "pamld", (8, 8), 40)}
    yield "dual index pamld", job, job
    job = {"sample": helpers.random_job(rng, "mdd", (6, 9), 32, minimum_distance=3),
           "molecular": [{"algorithm": "naive", "transform": {"token": ["0::4"]}}],
           "cellular": [helpers.random_job(rng, "pamld", (8,), 12, reverse=True), helpers.random_job(rng, "pamld", (5, 7), 20)]}
    job["cellular"][0]["transform"]["token"] = ["0:3:11"]
    job["cellular"][1]["transform"]["token"] = ["1:6:11", "0:1:8"]
    spelled = copy.deepcopy(job)
    spelled["cellular"][1]["transform"]["token"] = ["1:-5:", "0:1:8"]         # from the end of an 11 nt segment, open ended
    yield "chain", job, spelled
    knit = helpers.random_job(rng, "pamld", (10, 10), 30, **{"high quality threshold": 20, "high quality distance threshold": 2})
    knit["transform"] = {"token": ["0:0:6", "0:8:12", "1:2:12"], "knit": ["~0:1", "2"]}     # two tokens knit into one segment, the first reverse complemented
    job = {"cellular": [knit]}
    yield "knit", job, job


@pytest.mark.parametrize("short", [0.0, 0.25])
@pytest.mark.parametrize("sub_batch", [0, 777])
def test_raw_equals_host_packed(short, sub_batch, monkeypatch):
    if sub_batch:
        monkeypatch.setenv("PHQ_SUB_BATCH_READS", str(sub_batch))
    rng = np.random.default_rng(31 + sub_batch)
    for label, job, spelled in jobs(rng):
        n = 12000
        code, quality, offset, _ = workload.synthesize(compile_job(job), [0, 0], n, seed=3, short_fraction=short)
        compiled = compile_job(spelled)
        qcfail = (rng.random(n) < 0.1).astype(np.uint8)
        packed_chain = DecoderChain(compiled, device=0)
        tiles = packed_chain.pack(code, quality, offset)
        expected, expected_flags = packed_chain.decode(tiles, n, qcfail)

        sequence, ascii_quality = to_fastq_bytes(code, quality)
        raw_chain = DecoderChain(compiled, device=0)
        segments = [(sequence[i], ascii_quality[i], offset[i], 0) for i in range(len(code))]
        results, flags = raw_chain.decode_raw(segments, n, 33, qcfail)
        assert np.array_equal(flags, expected_flags), label
        for k in range(raw_chain.n_decoders):
            for field in ("index", "distance"):
                assert np.array_equal(results[k][field], expected[k][field]), (label, k, field)
            assert np.array_equal(results[k]["confidence"].view(np.uint64), expected[k]["confidence"].view(np.uint64)), (label, k)
            u, f = raw_chain.accumulators(k)
            eu, ef = packed_chain.accumulators(k)
            assert np.array_equal(u, eu), (label, k)
            assert np.allclose(f, ef, rtol=1e-12, atol=0), (label, k)
        assert raw_chain.totals() == packed_chain.totals()


def test_state_carries_over_calls_and_mixes_with_phq_pack():
    """Short PAMLD tokens read what earlier reads left in the Observation (barcode.h:150): three calls — raw, host
    packed, raw — give what one host packed call over everything gives."""
    rng = np.random.default_rng(8)
    job = {"sample": helpers.random_job(rng, "pamld", (8, 8), 24)}
    compiled = compile_job(job)
    n = 9000
    code, quality, offset, _ = workload.synthesize(compiled, [0, 0], n, seed=5, short_fraction=0.4)
    whole = DecoderChain(compiled, device=0)
    expected, expected_flags = whole.decode(whole.pack(code, quality, offset), n)

    sequence, ascii_quality = to_fastq_bytes(code, quality)
    chain = DecoderChain(compiled, device=0)
    cuts = [0, 2500, 6100, n]
    index, flags = [], []
    for part, (a, b) in enumerate(zip(cuts[:-1], cuts[1:])):
        part_offset = [o[a:b + 1] - o[a] for o in offset]
        part_code = [c[o[a]:o[b]] for c, o in zip(code, offset)]
        part_quality = [q[o[a]:o[b]] for q, o in zip(quality, offset)]
        if part == 1:
            results, f = chain.decode(chain.pack(part_code, part_quality, part_offset), b - a)
        else:
            segments = [(sequence[i][offset[i][a]:offset[i][b]], ascii_quality[i][offset[i][a]:offset[i][b]], part_offset[i], 0) for i in range(2)]
            results, f = chain.decode_raw(segments, b - a)
        index.append(results[0])
        flags.append(f)
    got = np.concatenate(index)
    assert np.array_equal(got["index"], expected[0]["index"])
    assert np.array_equal(got["confidence"].view(np.uint64), expected[0]["confidence"].view(np.uint64))
    assert np.array_equal(np.concatenate(flags), expected_flags)


def test_fixed_length_segments_and_compact_records():
    """Index reads of a run have one length: no offsets travel. Lower case and IUPAC letters decode as the feed does."""
    spec = workload.load("c1")
    compiled = compile_job(spec["job"])
    n = 50000
    code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], n, seed=17)
    sequence, ascii_quality = to_fastq_bytes(code, quality, phred_offset=64)
    sequence[0] = sequence[0].copy()
    lower = np.arange(sequence[0].shape[0]) % 7 == 0
    sequence[0][lower] = np.char.lower(sequence[0][lower].view("S1")).view(np.uint8)
    packed_chain = DecoderChain(compiled, device=0)
    expected = packed_chain.decode_compact(packed_chain.pack(code, quality, offset), n)
    raw_chain = DecoderChain(compiled, device=0)
    lengths = [int(o[1] - o[0]) for o in offset]
    assert all(np.all(np.diff(o) == l) for o, l in zip(offset, lengths))
    segments = [(sequence[i], ascii_quality[i], None, lengths[i]) for i in range(len(code))]
    results = raw_chain.decode_raw(segments, n, 64, compact=True)
    assert np.array_equal(results[0]["packed"], expected[0]["packed"])
    assert np.array_equal(results[0]["error_probability"].view(np.uint32), expected[0]["error_probability"].view(np.uint32))
    assert "pack" not in os.environ.get("PHQ_DISABLE", "")


@pytest.mark.parametrize("short", [0.0, 0.25])
@pytest.mark.parametrize("sub_batch", [0, 555])
def test_bam_segments_equal_the_oracle(short, sub_batch, monkeypatch):
    """phq_decode_batch_bam: the reference's own in-memory form of a segment (Sequence::code BAM bytes and Phred bytes,
    sequence.h:264-300) goes to the device unchanged; token slicing, reverse complement, knit and packing happen there.
    Checked against the oracle (the reference's classes when oracle/_ref is built) on the very same arrays."""
    if sub_batch:
        monkeypatch.setenv("PHQ_SUB_BATCH_READS", str(sub_batch))
    rng = np.random.default_rng(131 + sub_batch)
    for label, job, spelled in jobs(rng):
        n = 8000
        code, quality, offset, _ = workload.synthesize(compile_job(job), [0, 0], n, seed=23, short_fraction=short)
        compiled = compile_job(spelled)
        qcfail = (rng.random(n) < 0.1).astype(np.uint8)
        chain = DecoderChain(compiled, device=0)
        segments = [(code[i], quality[i], offset[i], 0) for i in range(len(code))]
        results, flags = chain.decode_raw(segments, n, 0, qcfail, bam=True)
        checker = O.best_oracle(compiled, len(code))
        expected = checker.decode(O.ReadBatch(code, quality, offset, qcfail))
        assert np.array_equal(flags, expected.qcfail), label
        for k, info in enumerate(chain.info):
            if info.algorithm == 0:
                helpers.compare_pamld(results[k], expected.index[:, k], expected.distance[:, k], expected.confidence[:, k], "%s decoder %d" % (label, k))
            elif info.has_tile:
                assert np.array_equal(results[k]["index"], expected.index[:, k]), (label, k)
                assert np.array_equal(results[k]["distance"], expected.distance[:, k]), (label, k)
            u, f = chain.accumulators(k)
            eu, ef = checker.accumulators(k)
            assert np.array_equal(u, eu), (label, k)
            assert np.allclose(f, ef, rtol=1e-9, atol=0), (label, k)
        assert chain.totals() == checker.totals()
        # and the compact records of the same call family
        again = DecoderChain(compiled, device=0)
        compact = again.decode_raw(segments, n, 0, qcfail, compact=True, bam=True)
        for k, info in enumerate(again.info):
            if info.has_tile:
                assert np.array_equal(compact[k]["packed"] & 0xffffff, results[k]["index"].astype(np.uint32)), (label, k)
