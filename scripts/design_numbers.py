"""Rewrites the measured table of DESIGN.md section 7 (between the BENCH_TABLE markers) from a bench JSON line.
usage: python scripts/design_numbers.py profiles/r02_bench_default.json profiles/r02_bench_reference.json"""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = json.loads(open(sys.argv[1]).read().splitlines()[0])
reference = json.loads(open(sys.argv[2]).read().splitlines()[0])
NOTES = {
    "c1": ("C1 PAMLD 96 × [8,8], 2^28 reads", "prefilter (separable) + exact + tie; issue slots"),
    "c2": ("C2 MDD 96 × [8,8], 2^26", "lookup kernel; issue slots"),
    "c3": ("C3 SPLiT-seq (4 PAMLD + naive), 2^24", "prefilter (generic, L = 8 / 6) + exact + tie per decoder; issue slots"),
    "c4": ("C4 sci-RNA-seq (2 PAMLD + naive), 2^24", "prefilter (generic, L = 10 / 20) + exact + tie; issue slots"),
    "c5": ("C5 whitelist 737,280 × [16] + naive, 1.25 × 10^8", "pruned bit-sliced scan (§4.9), one step of 13 s; issue slots"),
}


def fmt(x):
    return "%.3g" % x


def row(key, entry):
    name, note = NOTES[key]
    r = entry["roofline"]
    extra = ""
    if "two_pass" in entry:
        extra = " (pass 2, estimated priors: %s)" % fmt(entry["two_pass"]["pass2"]["value"])
    binding = r.get("binding", {})
    note = "%s %.0f %% (ncu)" % (note, binding.get("pct_of_peak", float("nan")))
    return "| %s | **%s**%s | %s ms | %.3f | %.2f | %s | %s |" % (
        name, fmt(entry["value"]), extra, fmt(r["kernel_ms_per_launch_set"]), r["frac"], r.get("int_equivalent", {}).get("frac", float("nan")),
        fmt(entry["cpu_baseline"]["value"]) if "cpu_baseline" in entry else "—", note)


rows = [row("c1", line)] + [row(k, line["configs"][k]) for k in ("c2", "c3", "c4", "c5") if k in line.get("configs", {})]
text = ("| workload | reads/s (HBM resident) | kernels per step | HBM roofline frac | INT-equivalent frac (§8d) | CPU reference reads/s (16 threads) | kernels, binding resource |\n"
        "|---|---|---|---|---|---|---|\n" + "\n".join(rows))


def e2e(key):
    v = line[key]
    return "%s reads/s (%d %% of the copy ceiling, %d + %d B/read)" % (
        fmt(v["value"]), round(100 * v["frac_of_copy_ceiling"]), v["h2d_bytes_per_step"] // v["reads_per_gpu_per_step"], v["d2h_bytes_per_step"] // v["reads_per_gpu_per_step"])


text += ("\n\nEnd to end on C1 (2^25 reads per call and rank, pinned host buffers): `e2e` (FASTQ bytes in) " + e2e("e2e") + ", `e2e_bam` " + e2e("e2e_bam")
         + ", `e2e_tags` " + e2e("e2e_tags") + ", `e2e_full` " + e2e("e2e_full") + ", `e2e_packed` " + e2e("e2e_packed")
         + ". The copy ceiling of the headline form is %.1f GB/s in and %.1f GB/s out per rank at once." % (
             line["e2e"]["copy_ceiling"]["per_rank_h2d_gbs"], line["e2e"]["copy_ceiling"]["per_rank_d2h_gbs"])
         + " Reference arm on the same box (`bench.py --impl reference`, `oracle/_ref`, %d threads): %s reads/s on C1, so the end-to-end form that starts from a feed's bytes is %d × the reference and the HBM-resident rate %d ×." % (
             reference["cpu_baseline"]["cores"], fmt(reference["value"]), round(line["e2e"]["value"] / reference["value"]), round(line["value"] / reference["value"])))
two = line["configs"]["c4"]["two_pass"]
one = line["two_pass"]
text += "\nThe two-pass workflow on one GPU: C4 pass 1 %.2f ms + finalize and install %.2f ms + pass 2 %.2f ms per 2^24 reads; C1 %.2f + %.2f + %.2f ms per 2^28 reads." % (
    two["pass1_ms"], two["finalize_and_install_ms"], two["pass2_ms"], one["pass1_ms"], one["finalize_and_install_ms"], one["pass2_ms"])

path = os.path.join(ROOT, "DESIGN.md")
design = open(path).read()
begin, end = "<!-- BENCH_TABLE_BEGIN -->", "<!-- BENCH_TABLE_END -->"
a, b = design.index(begin) + len(begin), design.index(end)
open(path, "w").write(design[:a] + "\n" + text + "\n" + design[b:])
print(text)
