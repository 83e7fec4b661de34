"""Markdown rows for DESIGN.md §7 from a bench.py JSON line. usage: bench_table.py bench.json"""
import json, sys
line = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][0])
def row(name, e):
    r = e["roofline"]
    extra = ""
    if "two_pass" in e:
        extra = "; pass 2 (estimated priors) %.3g" % e["two_pass"]["pass2"]["value"]
    cpu = e.get("cpu_baseline", {}).get("value")
    print("| %s | %.4g%s | %.4g ms | %.3g | %.2f | %s | %s |" % (name, e["value"], extra, r["kernel_ms_per_launch_set"], r["frac"], r.get("int_equivalent", {}).get("frac", float("nan")),
          ("%.3g" % cpu) if cpu else "—", r["kernel"].replace("|", "/")))
print("n_gpus", line["n_gpus"])
print("| config | reads/s (HBM resident) | kernels per step | HBM frac | INT-equivalent frac | CPU reference reads/s | kernels |")
print("|---|---|---|---|---|---|---|")
row("c1 (%d reads/GPU)" % line["config"]["reads_per_gpu"], line)
for k, v in line.get("configs", {}).items():
    if "error" not in v:
        row("%s (%d reads/GPU)" % (k, v["reads_per_gpu"]), v)
for k in ("e2e", "e2e_bam", "e2e_tags", "e2e_full", "e2e_packed"):
    if k in line:
        e = line[k]
        print("%s: %.4g reads/s, %.0f%% of the copy ceiling (%.1f GB/s in, %.1f GB/s out per rank), %d + %d bytes per read" % (
            k, e["value"], 100 * e["frac_of_copy_ceiling"], e["copy_ceiling"]["per_rank_h2d_gbs"], e["copy_ceiling"]["per_rank_d2h_gbs"],
            e["h2d_bytes_per_step"] // e["reads_per_gpu_per_step"], e["d2h_bytes_per_step"] // e["reads_per_gpu_per_step"]))
if "collect" in line:
    print("collect c1", line["collect"]["ms"], "ms", line["collect"]["bytes"], "bytes")
for k, v in line.get("configs", {}).items():
    if "collect" in v:
        print("collect", k, v["collect"]["ms"], "ms", v["collect"]["bytes"], "bytes", v.get("verified"))
    if "two_pass" in v:
        print("two_pass", k, {a: b for a, b in v["two_pass"].items() if a not in ("pass2", "kernels_pass2")})
if "two_pass" in line:
    print("two_pass c1", {a: b for a, b in line["two_pass"].items() if a not in ("pass2", "kernels_pass2")})
