"""Segment files on disk: `<key>_fst` (vellum v1 FST: term -> `_val` byte offset, or term -> the
single posting in direct mode) and `<key>_val` (the intcomp streams back to back, little-endian
words, no framing) — the format of file/writer.go and file/reader.go, so a directory written
here has the layout the Go index opens (shard.go:318-331) and vice versa.

  write_segment   Writer.Append x n + Close (file/writer.go:32-89): `_tmp` names first, renamed
                  when complete; no `_val` file in direct mode (NewDirectWriter, :95-121).
  open_segment    NewReader (file/reader.go:136-199): FST scoped to [min, max]; a missing `_val`
                  file switches to direct mode (:159-180); run i ends where run i+1 starts
                  (:52) or at the end of the file (:64).
  list_segments   the `<key>_fst` files of a shard directory, by key (shard.go:318-331).

The FST bytes come from libii2.so's host-side vellum restatement (csrc/fst_v1.cpp, bytes
unverified against Go); the `_val` bytes from the device encoder inside ii2_merge.
"""
from __future__ import annotations

import os

import numpy as np

from . import _abi as A
from . import fst as F
from .flat import FlatSegment, MergeResult


def write_segment(basedir: str, key: str, seg: FlatSegment | MergeResult) -> None:
    """`seg` is a VAL-mode or DIRECT-mode FlatSegment, or a MergeResult of ii2_merge."""
    if isinstance(seg, MergeResult):
        seg = seg.to_segment(key)
    if seg.mode == A.II2_SEG_DECODED:
        raise ValueError("encode the segment first (FlatSegment.to_val or ii2_merge)")
    data = F.fst_build(seg.term_bytes, seg.term_off, seg.val_off)
    fst_tmp = os.path.join(basedir, key + "_fst_tmp")
    with open(fst_tmp, "wb") as f:
        f.write(data)
    if seg.mode == A.II2_SEG_VAL:
        val_tmp = os.path.join(basedir, key + "_val_tmp")
        with open(val_tmp, "wb") as f:
            f.write(np.ascontiguousarray(seg.val_bytes, dtype=np.uint8)[: seg.val_size].tobytes())
    os.rename(fst_tmp, os.path.join(basedir, key + "_fst"))
    if seg.mode == A.II2_SEG_VAL:
        os.rename(os.path.join(basedir, key + "_val_tmp"), os.path.join(basedir, key + "_val"))


def open_segment(basedir: str, key: str, min_term: bytes | None = None,
                 max_term: bytes | None = None) -> FlatSegment | None:
    """The terms of the segment inside [min, max] as a flat view ready for ii2_merge /
    ii2_read_range; None when no term is in range (vellum.ErrIteratorDone, shard.go:257-261)."""
    with open(os.path.join(basedir, key + "_fst"), "rb") as f:
        data = f.read()
    # like the reader, do not bound the FST walk by max: the run of the last term in range ends
    # at the offset of the term after it (file/reader.go:144-146)
    tb, off, val, _ = F.fst_read(data, min_term, None)
    n = len(val)
    if n == 0:
        return None
    hi = n
    if max_term is not None:
        raw = tb.tobytes()
        lo_i, hi_i = 0, n  # first term > max
        while lo_i < hi_i:
            mid = (lo_i + hi_i) // 2
            if raw[int(off[mid]):int(off[mid + 1])] <= max_term:
                lo_i = mid + 1
            else:
                hi_i = mid
        hi = lo_i
        if hi == 0:
            return None
    val_path = os.path.join(basedir, key + "_val")
    t_off = off[: hi + 1].copy()
    t_bytes = tb[: int(t_off[-1])].copy()
    if not os.path.exists(val_path):  # direct mode
        return FlatSegment(t_bytes, t_off, A.II2_SEG_DIRECT, val_off=val[:hi].copy(), key=key)
    size = os.path.getsize(val_path)
    end = int(val[hi]) if hi < n else size
    start = int(val[0])
    with open(val_path, "rb") as f:
        f.seek(start)
        vb = np.frombuffer(f.read(end - start), dtype=np.uint8).copy()
    return FlatSegment(t_bytes, t_off, A.II2_SEG_VAL, val_bytes=vb,
                       val_off=(val[:hi] - np.uint64(start)).astype(np.uint64),
                       val_size=end - start, key=key)


def list_segments(basedir: str) -> list[str]:
    keys = [n[:-4] for n in os.listdir(basedir) if n.endswith("_fst")]
    return sorted(keys, key=lambda k: (len(k), k))


def remove_segment(basedir: str, key: str) -> None:
    for suffix in ("_fst", "_val"):
        try:
            os.remove(os.path.join(basedir, key + suffix))
        except FileNotFoundError:
            pass
