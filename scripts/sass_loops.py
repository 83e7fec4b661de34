"""Backward branches (loops) of one kernel in a cuobjdump -sass listing: body length and opcode histogram.
usage: cuobjdump -sass lib.so | python scripts/sass_loops.py KERNEL_SUBSTRING [minimum body length]"""
import re, sys, collections
want = sys.argv[1]; least = int(sys.argv[2]) if len(sys.argv) > 2 else 20
inside = False; code = []
for line in sys.stdin:
    if "Function :" in line:
        if inside: break
        inside = want in line
        continue
    if inside:
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m: code.append((int(m.group(1), 16), m.group(2)))
address = {a: i for i, (a, _) in enumerate(code)}
for i, (a, text) in enumerate(code):
    m = re.search(r"BRA(?:\.U(?:\.ANY)?)?\s+(?:!?U?P\d+,\s*)?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) <= a and int(m.group(1), 16) in address:
        first = address[int(m.group(1), 16)]
        if i - first + 1 < least: continue
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in code[first:i + 1])
        print("loop 0x%04x..0x%04x: %d instructions  %s" % (code[first][0], a, i - first + 1, " ".join("%s:%d" % kv for kv in ops.most_common(14))))
