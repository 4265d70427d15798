#!/usr/bin/env python3
"""One process, one C2 workload, many settings: per-kernel CUDA-event times of ii2_merge_dev.
usage: sweep2.py [--libs a.so,b.so] [--buckets 512,640,768] [--steps 6]
Every variant library is loaded side by side (different paths = different library instances)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from inverted_index_2_b200 import synth

ap = argparse.ArgumentParser()
ap.add_argument("--libs", default="")
ap.add_argument("--buckets", default="")
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--env", default="", help="extra settings swept per (lib, bucket): A=1;A=2+B=3")
ap.add_argument("--terms", type=int, default=1_000_000)
ap.add_argument("--segments", type=int, default=64)
ap.add_argument("--postings", type=int, default=100_000_000)
ap.add_argument("--presence", type=float, default=0.5)
ap.add_argument("--max-len", type=int, default=64)
a = ap.parse_args()
w = synth.make_workload(a.terms, a.segments, a.postings, removed_frac=0.05, presence=a.presence,
                        max_len=a.max_len)
# --libs "lib1:b1+b2,lib2:b3" (a lib named "default" = the in-tree build)
specs = [x for x in a.libs.split(",") if x] or ["default"]
for spec in specs:
    lib, _, bl = spec.partition(":")
    buckets = [x for x in bl.split("+") if x] or [x for x in a.buckets.split(",") if x] or [""]
    if lib == "default":
        lib = ""
    if lib:
        os.environ["II2_LIB"] = os.path.join(ROOT, lib)
    else:
        os.environ.pop("II2_LIB", None)
    from inverted_index_2_b200.engine import Engine
    eng = Engine(0)
    dsegs = [eng.upload(s) for s in w.segments]
    drem = eng.upload_removed(w.removed)
    eng.sync()
    envs = [x for x in a.env.split(";")] if a.env else [""]
    for bk, ev in [(b_, e_) for b_ in buckets for e_ in envs]:
        if bk:
            os.environ["II2_BUCKET"] = bk
        else:
            os.environ.pop("II2_BUCKET", None)
        for kv in envs:
            for one in kv.split("+"):
                if one:
                    os.environ.pop(one.split("=")[0], None)
        for one in ev.split("+"):
            if one:
                os.environ[one.split("=")[0]] = one.split("=")[1]
        ref = None
        for _ in range(3):
            r = eng.merge_dev(dsegs, drem, encode=True)
            info = r.info()
            sig = (int(info.terms_count), int(info.postings_out), int(info.val_size))
            r.release()
        eng.prof_enable(True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            eng.merge_dev(dsegs, drem, encode=True).release()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) / a.steps
        prof = eng.prof_read()
        eng.prof_enable(False)
        print(json.dumps({"lib": lib or "default", "bucket": bk or "default", "env": ev, "wall_ms": round(1e3 * wall, 3),
                          "sig": sig, **{p["name"]: round(p["ms"] / a.steps, 3) for p in prof}}), flush=True)
    for d in dsegs:
        d.release()
    drem.release()
    del eng
