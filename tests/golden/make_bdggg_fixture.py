"""Generate tests/golden/bdggg_* from the reference's own golden test (test/BDGGG, test/api/prior).

Run in the build container (needs /root/reference, read-only):

    python tests/golden/make_bdggg_fixture.py

Writes, next to this script:
  bdggg_reads.npz      the 250 x 3-segment input reads of BDGGG_s0{1,2,3}.fastq in the reference's in-memory
                       convention (BAM 4-bit code per base, Phred value per base, offsets, chastity qcfail)
  bdggg_job.json       the sample / molecular / cellular decoder directives of BDGGG_annotated.json with the
                       `base` decoder of BDGGG_interleave.json merged in (what the reference compiles)
  bdggg_expected.json  per output read of valid/annotated.out: name, flag and every tag (RG BC QT XB OX BZ CB CR CY XC) of the first
                       segment, with the order they are written in
  bdggg_report.json    valid/annotated.err (the JSON report with every accumulator and estimated prior)
  bdggg_compiled.json  the sample/molecular/cellular sections of valid/compile_annotated.out
  bdggg_import_documents.json  the job document of the annotated test and the document it imports, keyed by the
                       file names the import list uses (the tests write them out next to each other)
  prior_report.json / prior_estimated.json   test/api/prior input report and valid/BDGGG_annotated_estimated.json

Only data is written; no reference source is copied.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle.oracle import ReadBatch  # noqa: E402

REFERENCE = os.environ.get("PHENIQS_REFERENCE", "/root/reference")
T = os.path.join(REFERENCE, "test", "BDGGG")


def main():
    batch = ReadBatch.from_fastq([os.path.join(T, "BDGGG_s0%d.fastq" % i) for i in (1, 2, 3)])
    arrays = {"qcfail": batch.qcfail}
    for s in range(3):
        arrays["code%d" % s] = batch.code[s]
        arrays["quality%d" % s] = batch.quality[s]
        arrays["offset%d" % s] = batch.offset[s]
    with open(os.path.join(T, "BDGGG_s01.fastq"), "rb") as f:
        names = [l[1:].split(b" ")[0].decode() for l in f.read().split(b"\n")[0::4] if l]
    arrays["name"] = np.array(names)
    np.savez_compressed(os.path.join(HERE, "bdggg_reads.npz"), **arrays)

    annotated = json.load(open(os.path.join(T, "BDGGG_annotated.json")))
    interleave = json.load(open(os.path.join(T, "BDGGG_interleave.json")))

    def merge(base, overlay):
        out = json.loads(json.dumps(base))
        for k, v in overlay.items():
            if isinstance(v, dict) and isinstance(out.get(k), dict):
                out[k] = merge(out[k], v)
            else:
                out[k] = v
        return out

    def resolve(decoder):
        decoder = dict(decoder)
        base = decoder.pop("base", None)
        if base is not None:
            decoder = merge(interleave["decoder"][base], decoder)
        return decoder

    job = {
        "sample": resolve(annotated["sample"]),
        "molecular": [resolve(d) for d in annotated["molecular"]],
        "cellular": [resolve(d) for d in annotated["cellular"]],
        "min input length": annotated["min input length"],
        "filter incoming qc fail": interleave["filter incoming qc fail"],
    }
    json.dump(job, open(os.path.join(HERE, "bdggg_job.json"), "w"), indent=1, sort_keys=True)
    # the two job documents as the reference's test holds them (import + base inheritance; data only)
    json.dump({"BDGGG_annotated.json": annotated, "BDGGG_interleave.json": interleave},
              open(os.path.join(HERE, "bdggg_import_documents.json"), "w"), indent=1, sort_keys=True)

    expected = []
    for line in open(os.path.join(T, "valid", "annotated.out")):
        if line.startswith("@"):
            continue
        field = line.rstrip("\n").split("\t")
        flag = int(field[1])
        if not flag & 64:       # first segment only; the second repeats the tags (read.h:225-237)
            continue
        tag = {f[:2]: f[5:] for f in field[11:]}
        expected.append({"name": field[0], "flag": flag, "RG": tag.get("RG"), "BC": tag.get("BC"), "XB": tag.get("XB"),
                         "CB": tag.get("CB"), "CR": tag.get("CR"), "XC": tag.get("XC"), "OX": tag.get("OX"),
                         "QT": tag.get("QT"), "BZ": tag.get("BZ"), "CY": tag.get("CY"), "order": [f[:2] for f in field[11:]]})
    json.dump(expected, open(os.path.join(HERE, "bdggg_expected.json"), "w"), indent=0)

    report = json.load(open(os.path.join(T, "valid", "annotated.err")))
    json.dump(report, open(os.path.join(HERE, "bdggg_report.json"), "w"), indent=1, sort_keys=True)
    compiled = json.load(open(os.path.join(T, "valid", "compile_annotated.out")))
    json.dump({k: compiled[k] for k in ("sample", "molecular", "cellular")}, open(os.path.join(HERE, "bdggg_compiled.json"), "w"), indent=1, sort_keys=True)

    P = os.path.join(REFERENCE, "test", "api", "prior")
    json.dump(json.load(open(os.path.join(P, "BDGGG_annotated_report.json"))), open(os.path.join(HERE, "prior_report.json"), "w"), indent=1, sort_keys=True)
    json.dump(json.load(open(os.path.join(P, "valid", "BDGGG_annotated_estimated.json"))), open(os.path.join(HERE, "prior_estimated.json"), "w"), indent=1, sort_keys=True)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
