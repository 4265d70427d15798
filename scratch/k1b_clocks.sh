#!/bin/bash
# phase clocks of K1b (warp 0 of every CTA): builds with -DK1B_TIMING, runs one bench, prints shares
II2_NVCC_EXTRA="-DK1B_TIMING $1" python -m inverted_index_2_b200.build > /dev/null 2>&1 || { echo BUILD FAILED; exit 1; }
python - <<'PY'
import ctypes as C, sys, json, subprocess
sys.path.insert(0, ".")
import numpy as np
from inverted_index_2_b200 import synth
from inverted_index_2_b200.engine import Engine
eng = Engine(0)
w = synth.make_workload(1_000_000, 64, 100_000_000, seed=0xC2)
dsegs = [eng.upload(s) for s in w.segments]
drem = eng.upload_removed(w.removed)
for _ in range(3):
    eng.merge_dev(dsegs, drem).release()
out = (C.c_ulonglong * 10)()
eng.lib.ii2_debug_k1b_clocks(out, 1)
n = 5
for _ in range(n):
    eng.merge_dev(dsegs, drem).release()
eng.lib.ii2_debug_k1b_clocks(out, 0)
v = np.array(list(out), dtype=np.float64) / n
names = ["header", "tile choice", "run starts/reset", "search+offsets+keys", "hash", "wait keys/hash",
         "rank+records", "copies", "wait copies", "write-out"]
tot = v.sum()
for nm, x in zip(names, v):
    print(f"{nm:22s} {100 * x / tot:5.1f}%   {x / 41667:9.0f} clk per bucket")
print("total clk per bucket", tot / 41667)
PY
