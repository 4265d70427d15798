"""ctypes wrappers of the vellum v1 FST entry points (include/ii2.h: ii2_fst_*) — the
`<key>_fst` side of a segment (file/writer.go:35,43; file/reader.go:139-151).  Host-side code in
libii2.so; works without a GPU."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi as A

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        from .engine import load_library
        _LIB = load_library()
    return _LIB


class FstError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"ii2 fst: {what} (code {code})")
        self.code = code


def _buf(b: bytes | None):
    if b is None:
        return C.cast(None, A.u8p), 0, None
    keep = (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b).ljust(1, b"\0"))
    return C.cast(keep, A.u8p), len(b), keep


def fst_build(term_bytes: np.ndarray, term_off: np.ndarray, values: np.ndarray) -> bytes:
    """vellum.New + Insert(term i, values[i]) + Close over ascending distinct terms."""
    tb = np.ascontiguousarray(term_bytes, dtype=np.uint8)
    off = np.ascontiguousarray(term_off, dtype=np.uint32)
    val = np.ascontiguousarray(values, dtype=np.uint64)
    n = len(off) - 1
    assert len(val) == n
    if len(tb) == 0:
        tb = np.zeros(1, dtype=np.uint8)
    out, nb = A.u8p(), C.c_uint64()
    rc = _lib().ii2_fst_build(A.np_ptr(tb, A.u8p), A.np_ptr(off, A.u32p),
                              A.np_ptr(val, A.u64p) if n else C.cast(None, A.u64p), n,
                              C.byref(out), C.byref(nb))
    if rc != A.II2_OK:
        raise FstError(rc, "build")
    try:
        return C.string_at(out, nb.value)
    finally:
        _lib().ii2_fst_free(out)


def fst_build_items(items: list[tuple[bytes, int]]) -> bytes:
    terms = [t for t, _ in items]
    off = np.zeros(len(terms) + 1, dtype=np.uint32)
    if terms:
        off[1:] = np.cumsum([len(t) for t in terms])
    return fst_build(np.frombuffer(b"".join(terms), dtype=np.uint8), off,
                     np.array([v for _, v in items], dtype=np.uint64))


def fst_read(data: bytes, min_term: bytes | None = None, max_term: bytes | None = None):
    """(term_bytes, term_off, values, fst_len) of the keys in [min, max] (inclusive, None = open)."""
    p, n, k0 = _buf(data)
    pmin, nmin, k1 = _buf(min_term)
    pmax, nmax, k2 = _buf(max_term)
    out = A.FstTerms()
    rc = _lib().ii2_fst_read(p, n, pmin, nmin, pmax, nmax, C.byref(out))
    if rc != A.II2_OK:
        raise FstError(rc, "read")
    try:
        nt = int(out.n_terms)
        off = A.from_ptr(out.term_off, nt + 1, np.uint32)
        tb = A.from_ptr(out.term_bytes, int(off[-1]), np.uint8)
        val = A.from_ptr(out.values, nt, np.uint64)
        return tb, off, val, int(out.fst_len)
    finally:
        _lib().ii2_fst_terms_free(C.byref(out))


def fst_items(data: bytes, min_term: bytes | None = None, max_term: bytes | None = None
              ) -> list[tuple[bytes, int]]:
    tb, off, val, _ = fst_read(data, min_term, max_term)
    raw = tb.tobytes()
    return [(raw[int(off[i]):int(off[i + 1])], int(val[i])) for i in range(len(val))]


def fst_get(data: bytes, key: bytes) -> int | None:
    p, n, k0 = _buf(data)
    pk, nk, k1 = _buf(key)
    v, found = C.c_uint64(), C.c_int()
    rc = _lib().ii2_fst_get(p, n, pk, nk, C.byref(v), C.byref(found))
    if rc != A.II2_OK:
        raise FstError(rc, "get")
    return int(v.value) if found.value else None


def fst_len(data: bytes) -> int:
    p, n, k0 = _buf(data)
    v = C.c_uint64()
    rc = _lib().ii2_fst_len(p, n, C.byref(v))
    if rc != A.II2_OK:
        raise FstError(rc, "len")
    return int(v.value)


# ---- removed.list (gob stream of map[int64][]uint32, removed_list.go:26-33,73-80) -------------
def removed_list_encode(lists: dict[int, np.ndarray]) -> bytes:
    """RemovedLists.Serialize."""
    ts = np.array(sorted(lists), dtype=np.int64)
    parts = [np.ascontiguousarray(lists[int(t)], dtype=np.uint32) for t in ts]
    off = np.zeros(len(ts) + 1, dtype=np.uint64)
    if parts:
        off[1:] = np.cumsum([len(p) for p in parts])
    vals = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint32)
    if len(vals) == 0:
        vals = np.zeros(1, dtype=np.uint32)
    out, nb = A.u8p(), C.c_uint64()
    rc = _lib().ii2_removed_list_encode(
        ts.ctypes.data_as(C.POINTER(C.c_int64)) if len(ts) else C.cast(None, C.POINTER(C.c_int64)),
        A.np_ptr(off, A.u64p), A.np_ptr(vals, A.u32p), len(ts), C.byref(out), C.byref(nb))
    if rc != A.II2_OK:
        raise FstError(rc, "removed_list_encode")
    try:
        return C.string_at(out, nb.value)
    finally:
        _lib().ii2_fst_free(out)


def removed_list_decode(data: bytes) -> dict[int, np.ndarray]:
    """UnserializeRemovedList."""
    p, n, keep = _buf(data)
    out = A.RemovedLists()
    rc = _lib().ii2_removed_list_decode(p, n, C.byref(out))
    if rc != A.II2_OK:
        raise FstError(rc, "removed_list_decode")
    try:
        k = int(out.n_lists)
        ts = A.from_ptr(out.timestamps, k, np.int64)
        off = A.from_ptr(out.off, k + 1, np.uint64)
        vals = A.from_ptr(out.values, int(off[-1]) if k else 0, np.uint32)
        return {int(ts[i]): vals[int(off[i]):int(off[i + 1])].copy() for i in range(k)}
    finally:
        _lib().ii2_removed_lists_free(C.byref(out))
