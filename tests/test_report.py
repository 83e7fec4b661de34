"""The report writer and the prior adjusted job against the reference's stored outputs:
test/BDGGG/valid/annotated.err (the full report of the golden run) and
test/api/prior/valid/BDGGG_annotated_estimated.json (what the prior tool writes back)."""
import copy

import numpy as np
import pytest

import helpers
from oracle import oracle as O
from pheniqs_b200 import DecoderChain, adjust_job, compile_job


def bdggg_tables():
    batch, decoders, _ = helpers.bdggg()
    compiled = compile_job(decoders)
    oracle = O.PortOracle(compiled)
    oracle.decode(batch)
    tables = [oracle.accumulators(k) for k in range(3)]
    return compiled, tables, oracle.totals()


def test_report_reproduces_the_reference_report():
    """Every key and every digit (15 decimal places, truncated as rapidjson does) of annotated.err."""
    _, tables, totals = bdggg_tables()
    # the handle is built from the reference's own compile of the job (valid/compile_annotated.out), which carries
    # the read group tags the job-level compile projects onto the sample codec (transcode.cpp:1224-1262)
    chain = DecoderChain(helpers.golden("bdggg_compiled.json"), device=-1)      # host-only: no GPU needed
    golden = helpers.golden("bdggg_report.json")
    report = chain.encode_report(tables, totals, incoming=(golden["incoming"]["count"], golden["incoming"]["pf count"]))
    assert sorted(report) == sorted(golden)
    for section in golden:
        assert report[section] == golden[section], section
    # the text is key sorted at every level (json.cpp:875-893)
    text = chain.encode_report(tables, totals, text=True)
    keys = [line.strip().split('"')[1] for line in text.splitlines() if line.startswith('    "')]
    assert keys == sorted(keys) and "incoming" not in keys


def test_report_of_an_untouched_chain_is_clean():
    """No reads: 0 / 0 ratios are NaN in the reference and fail its `> 0` guards; zero counts stay, as there."""
    compiled, tables, _ = bdggg_tables()
    chain = DecoderChain(compiled, device=-1)
    empty = [(np.zeros_like(u), np.zeros_like(f)) for u, f in tables]
    report = chain.encode_report(empty, (0, 0))
    assert "outgoing" not in report and "estimated noise" not in report["sample"]
    assert report["sample"]["count"] == 0 and len(report["sample"]["classified"]) == 5
    assert all("estimated concentration" not in record for record in report["sample"]["classified"])


def test_adjust_job_reproduces_the_prior_tool():
    estimated = helpers.golden("prior_estimated.json")
    report = helpers.golden("prior_report.json")
    # the static job the tool started from: the stored result with the priors of test/api/prior/BDGGG_annotated.json
    original = {"@AGGCAGAA": 0.18, "@CGTACTAG": 0.20, "@GGACTCCT": 0.22, "@TAAGGCGA": 0.23, "@TCCTGAGC": 0.17}
    static = copy.deepcopy(estimated)
    static["sample"]["noise"] = 0.015
    for key, value in original.items():
        static["sample"]["codec"][key]["concentration"] = value
    adjusted = adjust_job(static, report)
    assert adjusted == estimated
    assert adjusted["sample"]["noise"] == report["sample"]["estimated noise"]
    # a barcode the report knows but has no estimate for is set to 0; an unknown one is left alone
    partial = copy.deepcopy(report)
    del partial["sample"]["classified"][0]["estimated concentration"]
    partial["sample"]["classified"][1]["barcode"] = ["TTTTTTTT"]
    again = adjust_job(static, partial)
    first = "@" + report["sample"]["classified"][0]["barcode"][0]
    second = "@" + report["sample"]["classified"][1]["barcode"][0]
    assert again["sample"]["codec"][first]["concentration"] == 0
    assert again["sample"]["codec"][second]["concentration"] == original[second]


def test_two_pass_workflow_on_the_host_side():
    """report -> adjust_job -> compile_job closes the loop of docs/pamld.md:38-44: the adjusted job compiles and its
    priors are the report's estimates, renormalised by the compile step (transcode.cpp:824-1039)."""
    _, decoders, _ = helpers.bdggg()
    compiled, tables, totals = bdggg_tables()
    chain = DecoderChain(compiled, device=-1)
    report = chain.encode_report(tables, totals)
    job = copy.deepcopy(decoders)
    for key, record in job["sample"]["codec"].items():
        record.setdefault("barcode", [key[1:]])
    adjusted = adjust_job(job, report)
    assert adjusted["sample"]["noise"] == report["sample"]["estimated noise"]
    again = compile_job(adjusted)
    estimates = np.array([r["estimated concentration"] for r in report["sample"]["classified"]])
    compiled_priors = np.array([again["sample"]["codec"][k]["concentration"] for k in sorted(again["sample"]["codec"])])
    expected = estimates / estimates.sum() * (1.0 - report["sample"]["estimated noise"])
    np.testing.assert_allclose(compiled_priors, expected, rtol=1e-12)


@pytest.mark.gpu
def test_report_from_the_device_accumulators():
    batch, decoders, _ = helpers.bdggg()
    compiled = compile_job(decoders)
    chain = DecoderChain(helpers.golden("bdggg_compiled.json"), device=0)
    for k in range(chain.n_decoders):       # the stored compile is truncated to 15 decimals: restore the exact priors
        section = compiled["sample"] if k == 0 else compiled["cellular"][0] if k == 2 else None
        if section is not None:
            chain.set_priors(k, section["noise"], [section["codec"][key]["concentration"] for key in sorted(section["codec"])])
    tiles = chain.pack(batch.code, batch.quality, batch.offset)
    chain.decode(tiles, batch.n_reads, batch.qcfail)
    golden = helpers.golden("bdggg_report.json")
    report = chain.report(incoming=(golden["incoming"]["count"], golden["incoming"]["pf count"]))

    def compare(a, b, path):
        assert type(a) is type(b) or (isinstance(a, (int, float)) and isinstance(b, (int, float))), path
        if isinstance(a, dict):
            assert sorted(a) == sorted(b), path
            for k in a:
                compare(a[k], b[k], path + "/" + k)
        elif isinstance(a, list):
            assert len(a) == len(b), path
            for i, (x, y) in enumerate(zip(a, b)):
                compare(x, y, "%s[%d]" % (path, i))
        elif isinstance(a, float):
            assert a == pytest.approx(b, rel=1e-12, abs=2e-15), path      # device sums are atomics: order differs
        else:
            assert a == b, path
    compare(report, golden, "")
