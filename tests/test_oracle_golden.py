"""The CPU oracles against the reference's own golden vectors (test/BDGGG, test/api/prior)."""
import numpy as np
import pytest

import helpers
from oracle import oracle as O


def oracles():
    kinds = [O.PortOracle]
    if O.ref_available():
        kinds.append(O.RefOracle)
    return kinds


@pytest.mark.parametrize("kind", oracles())
def test_bdggg_per_read_golden(kind):
    batch, decoders, expected = helpers.bdggg()
    compiled = O.compile_job(decoders)
    oracle = kind(compiled) if kind is O.PortOracle else kind(compiled, 3)
    out = oracle.decode(batch)
    rg = ["undetermined"] + [k[1:] for k in sorted(compiled["sample"]["codec"])]
    for i, e in enumerate(expected):
        assert (589 if out.qcfail[i] else 77) == e["flag"], e["name"]
        assert rg[out.index[i, 0]] == e["RG"].split(":")[-1], e["name"]
        assert helpers.error_tag(out.read_confidence[i, 0]) == e["XB"], e["name"]
        assert helpers.error_tag(out.read_confidence[i, 2]) == e["XC"], e["name"]
    # every PAMLD branch is present in the golden: pass, HQ-mismatch qcfail, low confidence, noise filter
    assert int(out.qcfail.sum()) == 12 and batch.n_reads == 248


@pytest.mark.parametrize("kind", oracles())
def test_bdggg_report_golden(kind):
    batch, decoders, _ = helpers.bdggg()
    compiled = O.compile_job(decoders)
    oracle = kind(compiled) if kind is O.PortOracle else kind(compiled, 3)
    oracle.decode(batch)
    report = helpers.golden("bdggg_report.json")
    # chain order: sample, molecular, cellular
    for k, section in ((0, report["sample"]), (2, report["cellular"][0])):
        u, f = oracle.accumulators(k)
        assert int(u[0, 0]) == section["unclassified"]["count"]
        assert int(u[0, 1]) == section["unclassified"]["pf count"]
        for row, record in enumerate(section["classified"], start=1):
            assert int(u[row, 0]) == record["count"]
            assert int(u[row, 1]) == record["pf count"]
            assert int(u[row, 3]) == record.get("low conditional confidence count", 0)
            assert int(u[row, 4]) == record.get("low confidence count", 0)
            if "average distance" in record:
                assert u[row, 2] / u[row, 0] == pytest.approx(record["average distance"], abs=2e-15)
            if "average confidence" in record:
                assert f[row, 0] / u[row, 0] == pytest.approx(record["average confidence"], abs=2e-15)
            if "average pf confidence" in record:
                assert f[row, 1] / u[row, 1] == pytest.approx(record["average pf confidence"], abs=2e-15)
        noise, concentration = oracle.estimate_priors(k)
        assert noise == pytest.approx(section["estimated noise"], abs=2e-15)
        for row, record in enumerate(section["classified"]):
            assert concentration[row] == pytest.approx(record["estimated concentration"], abs=2e-15)
    u, _ = oracle.accumulators(1)          # naive molecular decoder: everything lands on the undetermined row
    assert int(u[0, 0]) == 248
    count, pf_count = oracle.totals()
    assert (count, pf_count) == (report["incoming"]["count"] - 2, report["outgoing"]["pf count"]) or count == 248


def test_prior_api_golden():
    """test/api/prior: the estimated priors the reference's prior tool writes back into the configuration."""
    report = helpers.golden("prior_report.json")
    estimated = helpers.golden("prior_estimated.json")
    port = O.PortOracle(O.compile_job(helpers.bdggg()[1]))
    # the stored report is a 2,500 read run with a sample decoder only; its counts determine the priors
    for topic, section, target in (("sample", report["sample"], estimated["sample"]),):
        rows = len(section["classified"]) + 1
        u = np.zeros((rows, 6), dtype=np.uint64)
        f = np.zeros((rows, 2), dtype=np.float64)
        u[0, 0] = section["unclassified"]["count"]
        u[0, 1] = section["unclassified"]["pf count"]
        for row, record in enumerate(section["classified"], start=1):
            u[row, 0] = record["count"]
            u[row, 1] = record["pf count"]
            u[row, 3] = record.get("low conditional confidence count", 0)
            u[row, 4] = record.get("low confidence count", 0)
        noise, concentration = port.estimate_priors(0, (u, f))
        assert noise == pytest.approx(section["estimated noise"], abs=2e-15)
        assert noise == pytest.approx(target["noise"], abs=2e-15)
        keys = sorted(target["codec"])
        for row, key in enumerate(keys):
            assert concentration[row] == pytest.approx(target["codec"][key]["concentration"], abs=2e-15)


def test_compile_matches_reference_compile_output():
    """The restated decoder compile against valid/compile_annotated.out (printed at 15 decimals)."""
    _, decoders, _ = helpers.bdggg()
    compiled = O.compile_job(decoders)
    reference = helpers.golden("bdggg_compiled.json")
    for ours, theirs in ((compiled["sample"], reference["sample"]), (compiled["cellular"][0], reference["cellular"][0]), (compiled["molecular"][0], reference["molecular"][0])):
        for key in ("algorithm", "segment cardinality", "nucleotide cardinality", "barcode length", "confidence threshold", "noise",
                    "high quality threshold", "high quality distance threshold", "quality masking threshold", "index"):
            if key in theirs:
                assert ours[key] == theirs[key], key
        assert ours["transform"]["knit"] == theirs["transform"]["knit"]
        if "codec" in theirs:
            assert ours["distance tolerance"] == theirs["distance tolerance"]
            assert ours["random barcode probability"] == pytest.approx(theirs["random barcode probability"], abs=1e-15)
            for key, record in theirs["codec"].items():
                assert ours["codec"][key]["index"] == record["index"]
                assert ours["codec"][key]["concentration"] == pytest.approx(record["concentration"], abs=1.1e-15)


@pytest.mark.skipif(not O.ref_available(), reason="oracle/_ref not built")
def test_bdggg_tags_golden():
    """The checker of the tag path (RefOracle.tags: the reference's own Read::flush + Auxiliary members, read.h:187-237)
    reproduces every tag of test/BDGGG/valid/annotated.out, their order included, from the reference's own compile
    output — so the GPU tag tests compare against something that is itself pinned."""
    batch, _, expected = helpers.bdggg()
    compiled = helpers.golden("bdggg_compiled.json")
    tags, flags = O.RefOracle(compiled, 3).tags(batch)
    order = ("RG", "BC", "QT", "XB", "RX", "QX", "OX", "BZ", "XM", "CB", "CR", "CY", "XC")      # Auxiliary::encode, auxiliary.cpp:320-361
    for i, e in enumerate(expected):
        assert [t for t in order if t in tags[i]] == e["order"], e["name"]
        for tag in ("RG", "BC", "QT", "OX", "BZ", "CB", "CR", "CY"):
            assert tags[i].get(tag) == e[tag], "%s %s" % (e["name"], tag)
        for tag in ("XB", "XC"):
            assert (None if tag not in tags[i] else "%g" % tags[i][tag]) == e[tag], "%s %s" % (e["name"], tag)
        assert (589 if flags[i] else 77) == e["flag"], e["name"]
