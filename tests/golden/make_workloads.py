"""Generate pheniqs_b200/workloads/c{1..4}.json: the decoder directives of BASELINE.json's configs 1-4.

The barcode sets are the ones the reference ships as examples (data, not source):
  c1 / c2  decoder H7LT2DSXX_l03_sample of example/illumina_vignette/H7LT2DSXX_core.json (96 x [8,8]),
           PAMLD noise 0.05 / confidence threshold 0.95 (docs/illumina_vignette.md:137-139); c2 = same set, MDD tolerance [1,1]
  c3       example/splitseq_vignette/splitseq_l01_cellular.json (+ splitseq_core.json): 3 x 96 x [8] + 4 x [6] '~', UMI 2::10
  c4       example/scirnaseq_vignette/HGGKLBGX2_l01_cellular.json (+ HGGKLBGX2_core.json): 96 x [10] + 196 x [10,10], UMI 0::8
Config 5 (737,280 random 16-mers) has no example in the reference and is generated from a seed at run time
(pheniqs_b200/workload.py).

Run in the build container:  python tests/golden/make_workloads.py
"""
import json
import os

REFERENCE = os.environ.get("PHENIQS_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "..", "pheniqs_b200", "workloads")


def codec(core, name):
    return {key: {"barcode": value["barcode"]} for key, value in core["decoder"][name]["codec"].items()}


def resolve(directive, core):
    out = {k: v for k, v in directive.items() if k not in ("base", "comment")}
    if "base" in directive:
        out["codec"] = codec(core, directive["base"])
        for k, v in core["decoder"][directive["base"]].items():
            if k != "codec":
                out.setdefault(k, v)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    E = os.path.join(REFERENCE, "example")
    core = json.load(open(os.path.join(E, "illumina_vignette", "H7LT2DSXX_core.json")))
    base = core["decoder"]["H7LT2DSXX_l03_sample"]
    sample = {"algorithm": "pamld", "noise": 0.05, "confidence threshold": 0.95, "transform": base["transform"], "codec": codec(core, "H7LT2DSXX_l03_sample")}
    c1 = {"name": "c1", "description": "Illumina dual-index (i7+i5, 8 bp each) 96-sample PAMLD", "input segment length": [0, 8, 8, 0], "job": {"sample": sample}}
    mdd = dict(sample)
    mdd.update({"algorithm": "mdd", "distance tolerance": [1, 1], "quality masking threshold": 0})
    c2 = {"name": "c2", "description": "same 96-sample dual-index set, MDD, distance tolerance [1,1]", "input segment length": [0, 8, 8, 0], "job": {"sample": mdd}}

    core = json.load(open(os.path.join(E, "splitseq_vignette", "splitseq_core.json")))
    job = json.load(open(os.path.join(E, "splitseq_vignette", "splitseq_l01_cellular.json")))
    c3 = {"name": "c3", "description": "SPLiT-seq: 3 x 96 x [8] + 4 x [6] reverse complemented PAMLD cellular, naive 10 bp UMI", "input segment length": [0, 6, 94],
          "job": {"cellular": [resolve(d, core) for d in job["cellular"]], "molecular": [dict(resolve(d, core), algorithm="naive") for d in job["molecular"]]}}

    core = json.load(open(os.path.join(E, "scirnaseq_vignette", "HGGKLBGX2_core.json")))
    job = json.load(open(os.path.join(E, "scirnaseq_vignette", "HGGKLBGX2_l01_cellular.json")))
    c4 = {"name": "c4", "description": "sci-RNA-seq: 96 x [10] + 196 x [10,10] PAMLD cellular, naive 8 bp UMI, two-pass prior estimation", "input segment length": [18, 10, 10, 0],
          "job": {"cellular": [resolve(d, core) for d in job["cellular"]], "molecular": [dict(resolve(d, core), algorithm="naive") for d in job["molecular"]]}}

    for c in (c1, c2, c3, c4):
        json.dump(c, open(os.path.join(OUT, c["name"] + ".json"), "w"), indent=1, sort_keys=True)
        print(c["name"], {t: (len(v) if isinstance(v, list) else 1) for t, v in c["job"].items()})


if __name__ == "__main__":
    main()
