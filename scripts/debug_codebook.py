import numpy as np, sys
sys.path.insert(0, '.')
from pheniqs_b200 import DecoderChain, compile_job, workload
spec = workload.load("c1"); compiled = compile_job(spec["job"]); n = 50000
code, quality, offset, _ = workload.synthesize(compiled, spec["input segment length"], n, seed=13)
rng = np.random.default_rng(2)
quality9 = [np.array([2, 7, 11, 14, 22, 25, 30, 33, 37], dtype=np.uint8)[rng.integers(0, 9, size=q.shape)] for q in quality]
chain = DecoderChain(compiled, device=0)
wide = chain.pack(code, quality9, offset)
want, wf = chain.decode(wide, n)
four = DecoderChain(compiled, device=-1).pack(code, quality9, offset, quality_bits=-1)
print(four[0].quality_bits, list(four[0].quality_codebook))
a = workload.unpack_tile(wide[0].bases, wide[0].nmask, wide[0].quality, 16)
b = workload.unpack_tile(four[0].bases, four[0].nmask, four[0].quality, 16, four[0].quality_bits, four[0].quality_codebook)
print('unpack equal', np.array_equal(a[1], b[1]))
got, gf = chain.decode(four, n)
d = np.nonzero((got[0]['index'] != want[0]['index']) | (got[0]['confidence'] != want[0]['confidence']))[0]
print('diff reads', d.size, d[:10])
for r in d[:5]:
    print(r, want[0][r], got[0][r], a[1][r], a[0][r])
again, _ = chain.decode(wide, n)
print('byte form repeat equal', np.array_equal(again[0], want[0]))
