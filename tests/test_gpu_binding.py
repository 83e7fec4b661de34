"""The reference-side binding, compiled and run: oracle/batched_binding.cpp is the code a Pheniqs
maintainer would add — the reference's OWN Read / Segment / decoder classes, built against its headers and linked
with its unmodified objects, batching reads where TranscodingThread::run (transcode.h:202-225) classifies one, handing
the Segment buffers (BAM codes + Phred bytes) to phq_decode_batch_bam and scattering the verdicts back through the
reference's Read::append_to_* / update_* / set_RG and Read::flush (read.h:187-285).

Checked (a) against the reference's golden output test/BDGGG/valid/annotated.out (tests/golden/bdggg_expected.json:
flag, RG, BC, QT, XB, OX, BZ, CB, CR, CY, XC of all 248 reads, short index reads included) and the accumulators of
valid/annotated.err through the report, and (b) against the all-reference flow (RefOracle.tags: the same classes with
their own classify) on synthetic chains. Also runs the C++ wrapper's classify on the GPU (tests/cpp/wrapper_gpu.cpp)."""
import copy
import os
import subprocess

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from pheniqs_b200 import compile_job, workload

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(got, expected):
    return abs(float(got) - float(expected)) <= 1e-6 * float(expected) + 8 * 2.0 ** -53


def test_bdggg_through_the_reference_side_binding():
    if not O.binding_available():
        pytest.skip("oracle/_ref/libpheniqs_binding.so is not built")
    batch, decoders, expected = helpers.bdggg()
    compiled = helpers.golden("bdggg_compiled.json")        # the reference's own compile output (read group IDs included)
    for batch_reads in (2048, 100):                         # one feed batch, and three ragged ones
        tags, flags, report = O.batched_binding(compiled, batch, device=0, batch_reads=batch_reads)
        for r, e in enumerate(expected):
            got = tags[r]
            assert (589 if flags[r] else 77) == e["flag"], e["name"]
            for tag in ("RG", "BC", "QT", "OX", "BZ", "CB", "CR", "CY"):
                assert got.get(tag) == e[tag], "%s %s: %r, expected %r" % (e["name"], tag, got.get(tag), e[tag])
            for tag in ("XB", "XC"):
                assert (None if tag not in got else "%g" % got[tag]) == e[tag], "%s %s" % (e["name"], tag)
        golden = helpers.golden("bdggg_report.json")
        for key in ("count", "pf count", "classified count", "pf classified count", "low confidence count", "low conditional confidence count"):
            if key in golden["sample"]:
                assert report["sample"][key] == golden["sample"][key], key
        assert report["sample"]["estimated noise"] == pytest.approx(golden["sample"]["estimated noise"], abs=2e-15)


@pytest.mark.parametrize("short", [0.0, 0.25])
def test_binding_equals_the_all_reference_flow(short):
    if not (O.binding_available() and O.ref_available()):
        pytest.skip("oracle/_ref is not built")
    rng = np.random.default_rng(47)
    knit = helpers.random_job(rng, "pamld", (10, 10), 30, **{"high quality threshold": 20, "high quality distance threshold": 2})
    knit["transform"] = {"token": ["0:0:6", "0:8:12", "1:2:12"], "knit": ["~0:1", "2"]}
    job = {"sample": helpers.random_job(rng, "pamld", (8, 8), 40),
           "molecular": [helpers.random_job(rng, "pamld", (9,), 16, **{"corrected quality": 17}), {"algorithm": "naive", "transform": {"token": ["1:0:5"]}}],
           "cellular": [knit, helpers.random_job(rng, "mdd", (7,), 10, minimum_distance=3)]}
    job["molecular"][0]["transform"]["token"] = ["1:3:12"]
    job["cellular"][1]["transform"]["token"] = ["0:5:12"]
    compiled = compile_job(job)
    n = 5000
    code, quality, offset, _ = workload.synthesize(compiled, [0], n, seed=19, short_fraction=short)
    qcfail = (rng.random(n) < 0.1).astype(np.uint8)
    batch = O.ReadBatch(code, quality, offset, qcfail)
    expected, expected_flags = O.RefOracle(copy.deepcopy(compiled), len(code)).tags(batch)
    # one feed batch: with short reads the reference's decoders read what EARLIER reads left in their Observation, so
    # the batch boundaries of the binding must not matter either (the handle carries that state from call to call)
    for batch_reads in (n, 777):
        tags, flags, _ = O.batched_binding(compiled, batch, device=0, batch_reads=batch_reads)
        assert np.array_equal(flags, expected_flags)
        for r in range(n):
            assert sorted(tags[r]) == sorted(expected[r]), (r, tags[r], expected[r])
            for tag, value in expected[r].items():
                if tag in ("XB", "XM", "XC"):
                    assert close(tags[r][tag], value), (r, tag, tags[r][tag], value)
                else:
                    assert tags[r][tag] == value, (r, tag, tags[r][tag], value)


def test_cpp_wrapper_classifies_on_the_gpu(tmp_path):
    """include/pheniqs_b200.hpp driven from C++ on a device: compile, pack, classify, accumulators, report."""
    binary = str(tmp_path / "wrapper_gpu")
    library = os.path.join(ROOT, "pheniqs_b200")
    subprocess.run(["/usr/bin/g++", "-std=c++11", "-Wall", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "wrapper_gpu.cpp"),
                    "-L", library, "-lpheniqs_b200", "-Wl,-rpath," + library, "-o", binary], check=True)
    out = subprocess.run([binary], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.strip().endswith("ok"), out.stdout
