"""Python loader for the CPU oracle (oracle/libii2_oracle.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

from inverted_index_2_b200 import _abi as A  # noqa: E402
from inverted_index_2_b200.flat import (FlatSegment, MergeResult, ReadResult,  # noqa: E402
                                        views_array)

_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libii2_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in
            ("intcomp_ref.c", "roaring_ref.c", "merge_ref.c", "ii2_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "ii2.h"))
    if force or not os.path.exists(so) or any(
            os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        if os.path.exists(os.path.join(_HERE, "intcomp_ref.c")):
            subprocess.check_call(["make", "-C", _HERE, "-B", "libii2_oracle.so"],
                                  stdout=subprocess.DEVNULL)
    return so


_PROTOS = {
    "orc_intcomp_bound": (C.c_size_t, [C.c_size_t]),
    "orc_intcomp_encode": (C.c_size_t, [A.u32p, C.c_size_t, A.u32p]),
    "orc_intcomp_count": (C.c_size_t, [A.u32p, C.c_size_t]),
    "orc_intcomp_decode": (C.c_size_t, [A.u32p, C.c_size_t, A.u32p, C.c_size_t]),
    "orc_intcomp_encode_batch": (C.c_uint64, [A.u32p, A.u64p, C.c_uint64, A.u32p, A.u64p]),
    "orc_intcomp_count_batch": (C.c_uint64, [A.u32p, A.u64p, C.c_uint64, A.u64p]),
    "orc_intcomp_decode_batch": (C.c_int, [A.u32p, A.u64p, C.c_uint64, A.u64p, A.u32p]),
    "orc_bitmask_new": (C.c_void_p, [A.u32p, C.c_uint64]),
    "orc_bitmask_free": (None, [C.c_void_p]),
    "orc_bitmask_len": (C.c_uint64, [C.c_void_p]),
    "orc_bitmask_values": (A.u32p, [C.c_void_p]),
    "orc_bitmask_put": (C.c_int, [C.c_void_p, A.u32p, C.c_uint64, C.c_int, C.POINTER(A.u8p),
                                  C.POINTER(C.c_uint64)]),
    "orc_bitmask_get": (C.c_int, [C.c_void_p, A.u8p, C.c_uint64, C.POINTER(A.u32p),
                                  C.POINTER(C.c_uint64)]),
    "orc_merge": (C.c_int, [C.POINTER(A.SegView), C.c_int, A.u32p, C.c_uint64, C.c_uint32,
                            C.POINTER(A.MergeOut)]),
    "orc_merge_out_free": (None, [C.POINTER(A.MergeOut)]),
    "orc_read_range": (C.c_int, [C.POINTER(A.SegView), C.c_int, A.u8p, C.c_size_t, A.u8p,
                                 C.c_size_t, A.u32p, C.c_uint64, C.POINTER(A.ReadOut)]),
    "orc_read_out_free": (None, [C.POINTER(A.ReadOut)]),
    "orc_shard_key": (C.c_uint32, [A.u8p, C.c_size_t]),
    "orc_free": (None, [C.c_void_p]),
}


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        A.bind(_LIB, _PROTOS)
    return _LIB


class OracleError(RuntimeError):
    def __init__(self, code: int):
        super().__init__(f"oracle error {code}")
        self.code = code


def _bytes_arg(b: bytes | None):
    if b is None:
        return C.cast(None, A.u8p), 0, None
    buf = (C.c_uint8 * max(1, len(b))).from_buffer_copy(b.ljust(1, b"\0"))
    return C.cast(buf, A.u8p), len(b), buf


def _removed_arg(removed):
    if removed is None:
        return None, C.cast(None, A.u32p), 0
    r = np.ascontiguousarray(removed, dtype=np.uint32)
    if len(r) == 0:  # non-NULL pointer for "empty list given"
        keep = np.zeros(1, dtype=np.uint32)
        return keep, A.np_ptr(keep, A.u32p), 0
    return r, A.np_ptr(r, A.u32p), len(r)


# ---- codec -------------------------------------------------------------------
def intcomp_encode(values) -> np.ndarray:
    v = np.ascontiguousarray(values, dtype=np.uint32)
    out = np.zeros(lib().orc_intcomp_bound(len(v)), dtype=np.uint32)
    n = lib().orc_intcomp_encode(A.np_ptr(v, A.u32p), len(v), A.np_ptr(out, A.u32p))
    return out[:n].copy()


def intcomp_decode(words) -> np.ndarray:
    w = np.ascontiguousarray(words, dtype=np.uint32)
    n = lib().orc_intcomp_count(A.np_ptr(w, A.u32p), len(w))
    if n == C.c_size_t(-1).value:
        raise OracleError(A.II2_ERR_CORRUPT)
    out = np.zeros(max(1, n), dtype=np.uint32)
    got = lib().orc_intcomp_decode(A.np_ptr(w, A.u32p), len(w), A.np_ptr(out, A.u32p), n)
    if got != n:
        raise OracleError(A.II2_ERR_CORRUPT)
    return out[:n]


def intcomp_encode_batch(post: np.ndarray, post_off: np.ndarray):
    """One CompressUint32 call per list, concatenated (file/writer.go:49-56)."""
    post = np.ascontiguousarray(post, dtype=np.uint32)
    off = np.ascontiguousarray(post_off, dtype=np.uint64)
    n = len(off) - 1
    total = int(off[-1] - off[0]) if n else 0
    out = np.zeros(8 * n + 2 * total + 64, dtype=np.uint32)
    woff = np.zeros(n + 1, dtype=np.uint64)
    pos = lib().orc_intcomp_encode_batch(A.np_ptr(post, A.u32p), A.np_ptr(off, A.u64p), n,
                                         A.np_ptr(out, A.u32p), A.np_ptr(woff, A.u64p))
    return out[:pos].copy(), woff


def intcomp_decode_batch(words: np.ndarray, word_off: np.ndarray):
    words = np.ascontiguousarray(words, dtype=np.uint32)
    woff = np.ascontiguousarray(word_off, dtype=np.uint64)
    n = len(woff) - 1
    off = np.zeros(n + 1, dtype=np.uint64)
    total = lib().orc_intcomp_count_batch(A.np_ptr(words, A.u32p), A.np_ptr(woff, A.u64p), n,
                                          A.np_ptr(off, A.u64p))
    if total == C.c_uint64(-1).value:
        raise OracleError(A.II2_ERR_CORRUPT)
    out = np.zeros(max(1, total), dtype=np.uint32)
    rc = lib().orc_intcomp_decode_batch(A.np_ptr(words, A.u32p), A.np_ptr(woff, A.u64p), n,
                                        A.np_ptr(off, A.u64p), A.np_ptr(out, A.u32p))
    if rc != 0:
        raise OracleError(rc)
    return out[:total], off


# ---- merge / read --------------------------------------------------------------
def merge(segs: list[FlatSegment], removed=None, decoded: bool = True) -> MergeResult:
    arr = views_array(segs)
    keep, rp, nr = _removed_arg(removed)
    out = A.MergeOut()
    rc = lib().orc_merge(arr, len(segs), rp, nr, A.II2_MERGE_WANT_DECODED if decoded else 0,
                         C.byref(out))
    if rc != 0:
        raise OracleError(rc)
    try:
        return MergeResult.from_c(out, decoded)
    finally:
        lib().orc_merge_out_free(C.byref(out))


def read_range(segs: list[FlatSegment], min_term: bytes | None = None,
               max_term: bytes | None = None, removed=None) -> ReadResult:
    arr = views_array(segs)
    keep, rp, nr = _removed_arg(removed)
    pmin, nmin, k1 = _bytes_arg(min_term)
    pmax, nmax, k2 = _bytes_arg(max_term)
    out = A.ReadOut()
    rc = lib().orc_read_range(arr, len(segs), pmin, nmin, pmax, nmax, rp, nr, C.byref(out))
    if rc != 0:
        raise OracleError(rc)
    try:
        return ReadResult.from_c(out)
    finally:
        lib().orc_read_out_free(C.byref(out))


def ingest(docs: list[tuple[list[bytes], int]], removed=None, decoded: bool = True) -> MergeResult:
    """Shard.Put for every document (shard.go:33-67: terms sorted, one direct-mode segment whose
    every term carries the document's value; a term repeated inside a document is one FST key)
    followed by ONE Shard.Merge of those segments (shard.go:158-212)."""
    segs = [FlatSegment.direct(sorted(set(terms)), val) for terms, val in docs]
    return merge(segs, removed, decoded)


def prefix_search(segs: list[FlatSegment], prefixes: list[bytes]) -> dict[bytes, list[int]]:
    """The per-shard scan of InvertedIndex.PrefixSearch restated literally
    (inverted_index.go:196 sort; :251 Read(prefixes[0], nil); :266-271 stop past the greatest
    prefix; :274-279 HasPrefix test of every prefix, append; :289-292 sort + compact), over one
    shard's segments.  Pure-Python loop over the oracle's read: small cases only."""
    if not prefixes:
        return {}
    prefixes = sorted(prefixes)
    greatest = prefixes[-1]
    found: dict[bytes, list[int]] = {}
    for term, values in read_range(segs, prefixes[0], None).items():
        if greatest < term[:min(len(term), len(greatest))]:
            break
        for p in prefixes:
            if term.startswith(p):
                found.setdefault(p, []).extend(values)
    return {k: sorted(set(v)) for k, v in found.items()}


def shard_key(term: bytes) -> int:
    p, n, keep = _bytes_arg(term)
    return int(lib().orc_shard_key(p, n))


# ---- Bitmask --------------------------------------------------------------------
class Bitmask:
    """file/bitmask.go Bitmask[uint32] on the CPU oracle."""

    def __init__(self, init=None):
        v = np.ascontiguousarray(init if init is not None else [], dtype=np.uint32)
        self._h = lib().orc_bitmask_new(A.np_ptr(v, A.u32p) if len(v) else C.cast(None, A.u32p),
                                        len(v))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bitmask_free(self._h)
            self._h = None

    def all_values(self) -> np.ndarray:
        n = lib().orc_bitmask_len(self._h)
        return A.from_ptr(lib().orc_bitmask_values(self._h), n, np.uint32)

    def put(self, values, fast: bool = False) -> bytes:
        v = np.ascontiguousarray(values, dtype=np.uint32)
        bp, nb = A.u8p(), C.c_uint64()
        rc = lib().orc_bitmask_put(self._h, A.np_ptr(v, A.u32p), len(v), int(fast), C.byref(bp),
                                   C.byref(nb))
        if rc != 0:
            raise OracleError(rc)
        data = C.string_at(bp, nb.value)
        lib().orc_free(bp)
        return data

    def get(self, enc: bytes) -> np.ndarray:
        p, n, keep = _bytes_arg(enc)
        vp, vn = A.u32p(), C.c_uint64()
        rc = lib().orc_bitmask_get(self._h, p, n, C.byref(vp), C.byref(vn))
        if rc != 0:
            raise OracleError(rc)
        out = A.from_ptr(vp, vn.value, np.uint32)
        lib().orc_free(vp)
        return out
