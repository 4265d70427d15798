import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from inverted_index_2_b200 import synth
from inverted_index_2_b200.engine import Engine
eng = Engine(0)
w = synth.make_workload(1000000, 256, 100000000, seed=0xC3, presence=0.125)
dsegs = [eng.upload(s) for s in w.segments]
drem = eng.upload_removed(w.removed)
n = len(w.term_off) - 1
for frac in (0.001, 0.01, 0.1):
    span = max(1, int(n * frac)); lo = n // 3
    tlo = synth.term_at(w.term_bytes, w.term_off, lo); thi = synth.term_at(w.term_bytes, w.term_off, lo + span - 1)
    def call():
        r = eng.read_range_dev(dsegs, tlo, thi, drem); r.release()
    for _ in range(5): call()
    torch.cuda.synchronize()
    eng.prof_enable(True)
    t0 = time.perf_counter()
    for _ in range(20): call()
    dt = (time.perf_counter() - t0) / 20
    pr = eng.prof_read(); eng.prof_enable(False)
    print("frac", frac, "us/call (profiled)", round(dt * 1e6, 1),
          " ".join("%s=%.0f/%.0f" % (k['name'], 1e3 * k['ms'] / 20, 1e3 * k['host_ms'] / 20) for k in pr))
    ts = []
    for _ in range(20):
        t0 = time.perf_counter(); call(); ts.append(time.perf_counter() - t0)
    print("   unprofiled median us", round(1e6 * float(np.median(ts)), 1))
