#!/usr/bin/env python3
"""Small range reads over 256 resident segments (C3): median latency of `reps` reads per fraction.
Under ncu (launch list) the last read's kernels are the steady-state sequence."""
import argparse, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from inverted_index_2_b200 import synth
from inverted_index_2_b200.engine import Engine
ap = argparse.ArgumentParser()
ap.add_argument("--fracs", default="0.001,0.01")
ap.add_argument("--reps", type=int, default=50)
ap.add_argument("--terms", type=int, default=1_000_000)
ap.add_argument("--postings", type=int, default=100_000_000)
ap.add_argument("--prof", action="store_true")
ap.add_argument("--env", default="", help="A=1;A=0 settings swept")
a = ap.parse_args()
w = synth.make_workload(a.terms, 256, a.postings, seed=0xC3, presence=0.125)
eng = Engine(0)
dsegs = [eng.upload(s) for s in w.segments]
drem = eng.upload_removed(w.removed)
n = len(w.term_off) - 1
for env in (a.env.split(";") if a.env else [""]):
    for kv in [x for x in env.split("+") if x]:
        k, _, v = kv.partition("=")
        os.environ[k] = v
    for frac in [float(x) for x in a.fracs.split(",")]:
        rng = np.random.default_rng(3)
        span = max(1, int(n * frac))
        lat, sig = [], None
        for i in range(a.reps):
            lo = int(rng.integers(0, n - span + 1))
            tlo = synth.term_at(w.term_bytes, w.term_off, lo)
            thi = synth.term_at(w.term_bytes, w.term_off, lo + span - 1)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = eng.read_range_dev(dsegs, tlo, thi, drem)
            info = r.info()
            t1 = time.perf_counter()
            r.release()
            if i >= 5:
                lat.append(t1 - t0)
            sig = (int(info.terms_count), int(info.postings_in), int(info.postings_out))
        phases = host = None
        if a.prof:
            eng.prof_enable(True)
            r = eng.read_range_dev(dsegs, tlo, thi, drem)
            r.info()
            r.release()
            pr = eng.prof_read()
            phases = {p["name"]: round(1e3 * p["ms"] / max(1, p["count"]), 1) for p in pr}
            host = {p["name"]: round(1e3 * p["host_ms"] / max(1, p["count"]), 1) for p in pr}
            eng.prof_enable(False)
        print(json.dumps({"env": env, "frac": frac, "phase_us": phases, "host_us": host, "median_us": round(1e6 * float(np.median(lat)), 1),
                          "min_us": round(1e6 * float(np.min(lat)), 1),
                          "p90_us": round(1e6 * float(np.percentile(lat, 90)), 1), "last_sig": sig}), flush=True)
