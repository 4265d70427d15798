#!/usr/bin/env python3
"""bench_extra.py — the other BASELINE.json configurations, one JSON line each:
  C1  4 x Put(all terms) + Merge as one ii2_ingest call (device sort of every document)
  C3  term-range read union across 256 segments with 5 % removed_list filtering (µs per call)
  C4  file/bitmask + intcomp encode/decode sweep, posting-list lengths 16 .. 16M
  prefix  PrefixSearch batches over 64 resident segments (µs per call)
bench.py stays the headline (C2 compaction); these lines are evidence for SURVEY §8 rows."""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from inverted_index_2_b200 import synth  # noqa: E402


def _peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


PEAK = _peak()  # GB/s, measured copy bandwidth (fallback: B200_PROFILING.md)


def timed(fn, reps):
    import torch
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts)), float(np.percentile(ts, 99))


def c3(eng, a):
    w = synth.make_workload(a.terms, 256, a.postings, seed=0xC3, presence=0.125)
    dsegs = [eng.upload(s) for s in w.segments]
    drem = eng.upload_removed(w.removed)
    n = len(w.term_off) - 1
    rng = np.random.default_rng(3)
    out = []
    for frac in (0.001, 0.01, 0.1, 1.0):
        span = max(1, int(n * frac))
        lat, info = [], None
        for _ in range(16 if frac < 1 else 4):
            lo = int(rng.integers(0, n - span + 1))
            tlo = synth.term_at(w.term_bytes, w.term_off, lo)
            thi = synth.term_at(w.term_bytes, w.term_off, lo + span - 1)

            def call():
                nonlocal info
                r = eng.read_range_dev(dsegs, tlo, thi, drem)
                info = r.info()
                r.release()
            med, _ = timed(call, 3)
            lat.append(med)
        eng.prof_enable(True)
        call()
        phases = {p["name"]: round(1e3 * p["ms"] / max(1, p["count"]), 1) for p in eng.prof_read()}
        host = {p["name"]: round(1e3 * p["host_ms"] / max(1, p["count"]), 1) for p in eng.prof_read()}
        eng.prof_enable(False)
        out.append({"range_frac": frac, "terms": int(info.terms_count), "phase_us": phases,
                    "host_us": host,
                    "postings_in": int(info.postings_in), "postings_out": int(info.postings_out),
                    "median_us": 1e6 * float(np.median(lat)), "p99_us": 1e6 * float(np.max(lat)),
                    "postings_in_per_s": int(info.postings_in) / float(np.median(lat))})
    print(json.dumps({"config": "C3 range read, 256 segments, 5% removed (read + merge-style filter)",
                      "terms": a.terms, "postings": w.postings_in, "results": out}))


def prefix(eng, a):
    """PrefixSearch (inverted_index.go:192-295) as one ii2_prefix_search_dev call over resident
    segments: batches of 16 prefixes of 3, 2 and 1 bytes, and the empty prefix (everything)."""
    w = synth.make_workload(a.terms, 64, a.postings, seed=0xC2)
    dsegs = [eng.upload(s) for s in w.segments]
    n = len(w.term_off) - 1
    rng = np.random.default_rng(6)
    out = []
    for plen, count in ((3, 16), (2, 16), (1, 16), (0, 1)):
        picks = sorted({synth.term_at(w.term_bytes, w.term_off, int(i))[:plen]
                        for i in rng.integers(0, n, size=count)})
        res = None

        def call():
            nonlocal res
            res = eng.prefix_search_dev(dsegs, picks)
        med, worst = timed(call, 3)
        eng.prof_enable(True)
        call()
        phases = {p["name"]: round(1e3 * p["ms"] / max(1, p["count"]), 1) for p in eng.prof_read()}
        eng.prof_enable(False)
        out.append({"prefix_len": plen, "prefixes": len(picks),
                    "values_out": int(sum(len(v) for v in res.values())),
                    "median_us": 1e6 * med, "max_us": 1e6 * worst, "phase_us": phases})
    print(json.dumps({"config": "PrefixSearch over 64 resident segments (one C-ABI call per batch, "
                                "results downloaded)", "terms": a.terms, "postings": w.postings_in,
                      "results": out}))


def c1(eng, a):
    """BASELINE configs[0] shape through ii2_ingest: 4 x Put(all terms, val = 1..4) + Merge as
    ONE call — every document's terms are handed over SHUFFLED (Put sorts them, shard.go:34)."""
    tb, off = synth.make_terms(a.terms)
    raw = tb.tobytes()
    terms = [raw[int(off[i]):int(off[i + 1])] for i in range(a.terms)]
    rng = np.random.default_rng(1)
    docs = []
    for val in (1, 2, 3, 4):
        perm = rng.permutation(a.terms)
        docs.append(([terms[int(i)] for i in perm], val))
    res = None

    def call():
        nonlocal res
        res = eng.ingest(docs, None, decoded=True)
    t0 = time.perf_counter()
    call()
    first = time.perf_counter() - t0
    eng.prof_enable(True)
    t0 = time.perf_counter()
    call()
    second = time.perf_counter() - t0
    phases = {p["name"]: round(p["ms"] / max(1, p["count"]), 2) for p in eng.prof_read()}
    eng.prof_enable(False)
    ok = (res.terms_count == a.terms and np.array_equal(res.term_bytes, tb)
          and np.array_equal(res.post, np.tile(np.array([1, 2, 3, 4], dtype=np.uint32), a.terms)))
    print(json.dumps({"config": "C1 ingest: 4 documents x all terms (shuffled), one ii2_ingest call "
                                "(ctypes marshalling of 4 x %d Python terms included)" % a.terms,
                      "terms": a.terms, "postings": 4 * a.terms, "first_call_ms": 1e3 * first,
                      "call_ms": 1e3 * second, "device_phase_ms": phases,
                      "result_is_every_term_to_1234": bool(ok)}))


def c4(eng, a):
    from oracle import orc  # checker only
    rows = []
    rng = np.random.default_rng(4)
    for lg in range(4, 25, 4):
        L = 1 << lg
        nlists = max(1, a.c4_values // L)  # >= 64M values per timing
        for gap in (1, 16, 4096):
            gaps = rng.integers(1, 2 * gap + 1, size=L * nlists, dtype=np.int64)
            starts = np.arange(nlists, dtype=np.int64) * L
            c = np.cumsum(gaps)
            vals = (c - np.repeat(c[starts] - gaps[starts], L)).astype(np.uint32)
            off = (np.arange(nlists + 1, dtype=np.uint64) * np.uint64(L))
            words = woff = None

            words, woff = eng.intcomp_encode_batch(vals, off)
            # the C-ABI calls themselves, on pinned host buffers (no numpy copies of the results)
            import ctypes as C
            import torch
            from inverted_index_2_b200 import _abi as A
            keep = [torch.from_numpy(x).pin_memory() for x in (vals, off, words, woff)]
            pv, po, pw, pwo = (t.numpy() for t in keep)

            def enc():
                wp, op = A.u32p(), A.u64p()
                eng._check(eng.lib.ii2_intcomp_encode_u32(A.np_ptr(pv, A.u32p), A.np_ptr(po, A.u64p),
                                                          nlists, C.byref(wp), C.byref(op)), "enc")
                eng.lib.ii2_free(wp)
                eng.lib.ii2_free(op)

            def dec():
                vp, op = A.u32p(), A.u64p()
                eng._check(eng.lib.ii2_intcomp_decode_u32(A.np_ptr(pw, A.u32p), A.np_ptr(pwo, A.u64p),
                                                          nlists, C.byref(vp), C.byref(op)), "dec")
                eng.lib.ii2_free(vp)
                eng.lib.ii2_free(op)
            te, _ = timed(enc, 3)
            td, _ = timed(dec, 3)
            eng.prof_enable(True)
            for _ in range(3):
                enc()
                dec()
            ph = {k["name"]: k["ms"] / k["count"] for k in eng.prof_read()}
            eng.prof_enable(False)
            alg = 4.0 * len(vals) + 4.0 * len(words)  # decoded values + stream words, each once
            if lg <= 12:  # parity spot check on the small cases
                ew, eo = orc.intcomp_encode_batch(vals[:L * min(nlists, 64)], off[:min(nlists, 64) + 1])
                assert np.array_equal(ew, words[:len(ew)]) and np.array_equal(eo, woff[:len(eo)])
            rows.append({"codec": "intcomp", "L": L, "lists": nlists, "gap": gap,
                         "ratio": 4.0 * len(vals) / (4.0 * len(words)),
                         "encode_values_per_s": len(vals) / te, "decode_values_per_s": len(vals) / td,
                         "api": "ii2_intcomp_*_u32 on pinned host buffers (H2D + kernels + D2H)",
                         "device_encode_ms": ph.get("k3a_encode"), "device_decode_ms": ph.get("k3a_decode"),
                         "device_encode_gbs": alg / ph["k3a_encode"] / 1e6 if ph.get("k3a_encode") else None,
                         "device_decode_gbs": alg / ph["k3a_decode"] / 1e6 if ph.get("k3a_decode") else None,
                         "encode_frac_of_hbm_peak": alg / ph["k3a_encode"] / 1e6 / PEAK if ph.get("k3a_encode") else None,
                         "decode_frac_of_hbm_peak": alg / ph["k3a_decode"] / 1e6 / PEAK if ph.get("k3a_decode") else None})
    for lg in range(4, 25, 4):
        L = 1 << lg
        universe = np.sort(rng.choice(max(4 * L, 1 << 10), size=2 * L, replace=False)).astype(np.uint32)
        vals = rng.permutation(universe)[:L]
        bm = eng.bitmask(universe)
        enc = None

        enc = bm.put(vals)
        import ctypes as C
        import torch
        from inverted_index_2_b200 import _abi as A
        tv = torch.from_numpy(np.ascontiguousarray(vals)).pin_memory()
        te_ = torch.from_numpy(np.frombuffer(enc, dtype=np.uint8).copy()).pin_memory()
        pv, pe = tv.numpy(), te_.numpy()

        def put():  # the C-ABI call itself on pinned buffers (no Python copies of the result)
            bp, nb = A.u8p(), C.c_uint64()
            eng._check(eng.lib.ii2_bitmask_put(bm.h, A.np_ptr(pv, A.u32p), L, C.byref(bp), C.byref(nb)), "put")
            eng.lib.ii2_free(bp)

        def get():
            vp, vn = A.u32p(), C.c_uint64()
            eng._check(eng.lib.ii2_bitmask_get(bm.h, A.np_ptr(pe, A.u8p), len(pe), C.byref(vp), C.byref(vn)), "get")
            eng.lib.ii2_free(vp)
        tp, _ = timed(put, 3)
        tg, _ = timed(get, 3)
        eng.prof_enable(True)
        for _ in range(3):
            put()
            get()
        ph = {k["name"]: k["ms"] / k["count"] for k in eng.prof_read()}
        eng.prof_enable(False)
        alg = 4.0 * L + len(enc)  # values + serialised bitmap, each once
        rows.append({"codec": "bitmask", "L": L, "dictionary": 2 * L, "bytes": len(enc),
                     "put_values_per_s": L / tp, "get_values_per_s": L / tg,
                     "api": "ii2_bitmask_put/get on pinned host buffers (H2D + kernels + D2H)",
                     "device_put_ms": ph.get("k3b_put"), "device_get_ms": ph.get("k3b_get"),
                     "note": "device_*_ms = the call's stream time INCLUDING its H2D / D2H copies "
                             "(64 MB of values at L = 16 M); kernels_*_ms = kernels only",
                     "kernels_put_ms": ph.get("k3b_put_kernels"), "kernels_get_ms": ph.get("k3b_get_kernels"),
                     "kernels_put_values_per_s": L / (ph["k3b_put_kernels"] * 1e-3) if ph.get("k3b_put_kernels") else None,
                     "kernels_get_values_per_s": L / (ph["k3b_get_kernels"] * 1e-3) if ph.get("k3b_get_kernels") else None,
                     "kernels_put_gbs": alg / ph["k3b_put_kernels"] / 1e6 if ph.get("k3b_put_kernels") else None,
                     "device_put_values_per_s": L / (ph["k3b_put"] * 1e-3) if ph.get("k3b_put") else None,
                     "device_put_gbs": alg / ph["k3b_put"] / 1e6 if ph.get("k3b_put") else None,
                     "device_get_gbs": alg / ph["k3b_get"] / 1e6 if ph.get("k3b_get") else None})
    print(json.dumps({"config": "C4 codec sweep, list lengths 16..16M", "results": rows}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="c3,c4")
    ap.add_argument("--c4-values", type=int, default=1 << 26)
    ap.add_argument("--terms", type=int, default=1_000_000)
    ap.add_argument("--postings", type=int, default=100_000_000)
    a = ap.parse_args()
    from inverted_index_2_b200.engine import Engine
    eng = Engine(0)
    if "c1" in a.which:
        c1(eng, a)
    if "c3" in a.which:
        c3(eng, a)
    if "c4" in a.which:
        c4(eng, a)
    if "prefix" in a.which:
        prefix(eng, a)


if __name__ == "__main__":
    main()
