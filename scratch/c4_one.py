"""One 16 M-value list through ii2_intcomp_encode_u32 / _decode_u32 (the top of the C4 sweep):
the ncu target for the long-list codec kernels."""
import sys
sys.path.insert(0, '/root/repo')
import numpy as np
from inverted_index_2_b200.engine import Engine
eng = Engine(0)
rng = np.random.default_rng(4)
L = 1 << 24
vals = np.cumsum(rng.integers(1, 33, size=L, dtype=np.int64)).astype(np.uint32)
off = np.array([0, L], dtype=np.uint64)
for _ in range(2):
    words, woff = eng.intcomp_encode_batch(vals, off)
    dec, doff = eng.intcomp_decode_batch(words, woff)
assert np.array_equal(dec, vals)
print("ok", len(words))
