"""Per kernel counts of the SASS mnemonics that show the hardware paths the design relies on (cuobjdump -sass of the built library).
usage: sass_evidence.py > profiles/r02_sass_evidence.txt"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pheniqs_b200", "libpheniqs_b200.so")
COLUMNS = ("UBLKCP", "SYNCS", "LDS", "STS", "LOP3", "POPC", "MATCH", "VOTE", "REDUX", "SHFL", "FMUL", "FMNMX", "DMUL", "DADD", "DFMA", "ATOMS", "ATOMG", "RED", "HMMA", "UTCHMMA")
out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
kernels, name = {}, None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(phq::DecoderParams.*", "", name).replace("phq::(anonymous namespace)::", "").replace("phq::", "").replace("(int)", "").replace("(bool)", "")
        kernels[name] = {c: 0 for c in COLUMNS}
        kernels[name]["total"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        op = m.group(1)
        kernels[name]["total"] += 1
        for c in COLUMNS:
            if op.startswith(c):
                kernels[name][c] += 1
print("SASS evidence per kernel (cuobjdump -sass pheniqs_b200/libpheniqs_b200.so, sm_100a): instruction counts of the mnemonics that")
print("show the hardware paths the design relies on. UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier arrive / try_wait,")
print("LOP3 = three-input logic (mismatch masks, carry-save adders), MATCH/VOTE/REDUX/SHFL = warp cooperation, FMUL/FMNMX = the f32")
print("prefilter scans, DMUL/DADD/DFMA = f64 pipe, ATOMS = shared-memory atomics (per-CTA accumulators; ATOMS.CAST = compare-and-swap")
print("loops are gone from the confidence sums), POPC. No HMMA / UTCMMA: the path is gather-and-compare, not a contraction.\n")
print("%-62s" % "kernel" + "".join("%8s" % c for c in COLUMNS) + "   total")
for name in sorted(kernels):
    print("%-62s" % name[:62] + "".join("%8d" % kernels[name][c] for c in COLUMNS) + "%8d" % kernels[name]["total"])
