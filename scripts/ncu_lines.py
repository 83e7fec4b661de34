"""Per source line: warp instructions executed and stall samples of one kernel of an .ncu-rep (needs -lineinfo and --import-source on).
usage: ncu_lines.py report.ncu-rep [kernel index (1-based)] [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]; kid = sys.argv[2] if len(sys.argv) > 2 else "1"; top = int(sys.argv[3]) if len(sys.argv) > 3 else 45
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
header = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
names = rows[header]
ix = names.index("Instructions Executed"); sx = names.index("# Samples")
lines = []
for r in rows[header + 1:]:
    if len(r) <= ix or not r[0].isdigit():
        continue
    try:
        lines.append((int(r[ix]), int(r[sx]), int(r[0]), r[1].strip()))
    except ValueError:
        pass
total = sum(l[0] for l in lines); samples = sum(l[1] for l in lines)
print(rows[1][1][:120]); print("total warp instructions", total, "samples", samples)
for n, s, line, text in sorted(lines, reverse=True)[:top]:
    print("%5.1f%% inst %5.1f%% stall  line %4d  %s" % (100.0 * n / total, 100.0 * s / max(samples, 1), line, text[:140]))
