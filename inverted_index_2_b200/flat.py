"""Flat (SoA) host representation of segments and results.

A FlatSegment is what the Go side hands over after iterating one segment's FST
once: concatenated term bytes + offsets, and the postings in one of the three
forms of include/ii2.h (decoded lists, raw `_val` bytes + FST outputs, or direct
mode).  It replaces the per-term file.TermValues objects (file/types.go:9-12).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from . import _abi as A


@dataclass
class FlatSegment:
    term_bytes: np.ndarray  # uint8
    term_off: np.ndarray  # uint32 [n+1]
    mode: int = A.II2_SEG_DECODED
    post: np.ndarray | None = None  # uint32
    post_off: np.ndarray | None = None  # uint64 [n+1]
    val_bytes: np.ndarray | None = None  # uint8
    val_off: np.ndarray | None = None  # uint64 [n]
    val_size: int = 0
    key: str = ""
    val_woff32: np.ndarray | None = None  # uint32 [n]: val_off / 4 (optional, `_val` views)

    @property
    def n_terms(self) -> int:
        return len(self.term_off) - 1

    def term(self, i: int) -> bytes:
        return self.term_bytes[self.term_off[i]:self.term_off[i + 1]].tobytes()

    def terms(self) -> list[bytes]:
        tb = self.term_bytes.tobytes()
        off = self.term_off
        return [tb[off[i]:off[i + 1]] for i in range(self.n_terms)]

    def view(self) -> A.SegView:
        """ctypes view; borrows this object's arrays (keep `self` alive)."""
        v = A.SegView()
        v.n_terms = self.n_terms
        v.term_bytes = A.np_ptr(self.term_bytes, A.u8p)
        v.term_off = A.np_ptr(self.term_off, A.u32p)
        v.mode = self.mode
        v.post = A.np_ptr(self.post, A.u32p)
        v.post_off = A.np_ptr(self.post_off, A.u64p)
        v.val_bytes = A.np_ptr(self.val_bytes, A.u8p)
        v.val_off = A.np_ptr(self.val_off, A.u64p)
        v.val_size = int(self.val_size)
        v.val_woff32 = A.np_ptr(self.val_woff32, A.u32p)
        return v

    # ---- constructors -------------------------------------------------------
    @staticmethod
    def _pack_terms(terms: list[bytes]):
        lens = np.fromiter((len(t) for t in terms), dtype=np.uint32, count=len(terms))
        off = np.zeros(len(terms) + 1, dtype=np.uint32)
        np.cumsum(lens, out=off[1:])
        tb = np.frombuffer(b"".join(terms), dtype=np.uint8).copy()
        return tb, off

    @classmethod
    def from_items(cls, items, key: str = "") -> "FlatSegment":
        """items: iterable of (term bytes, list of uint32) in ascending term order
        (values kept in the given order — the codec accepts unsorted lists,
        file/writer_test.go:14)."""
        items = list(items)
        tb, off = cls._pack_terms([t for t, _ in items])
        lens = np.fromiter((len(v) for _, v in items), dtype=np.uint64, count=len(items))
        poff = np.zeros(len(items) + 1, dtype=np.uint64)
        np.cumsum(lens, out=poff[1:])
        post = np.fromiter((x for _, v in items for x in v), dtype=np.uint32, count=int(poff[-1]))
        return cls(tb, off, A.II2_SEG_DECODED, post=post, post_off=poff, key=key)

    @classmethod
    def direct(cls, terms: list[bytes], val: int, key: str = "") -> "FlatSegment":
        """Direct-mode segment as Shard.Put writes it (shard.go:33-67): every term
        carries the single value `val` as its FST output (file/writer.go:34-40)."""
        terms = sorted(terms)
        tb, off = cls._pack_terms(terms)
        voff = np.full(len(terms), val, dtype=np.uint64)
        return cls(tb, off, A.II2_SEG_DIRECT, val_off=voff, key=key)

    def lists(self) -> list[list[int]]:
        if self.mode == A.II2_SEG_DIRECT:
            return [[int(v) & 0xFFFFFFFF] for v in self.val_off]
        assert self.mode == A.II2_SEG_DECODED
        return [self.post[int(self.post_off[i]):int(self.post_off[i + 1])].tolist()
                for i in range(self.n_terms)]

    def to_val(self, encode_batch) -> "FlatSegment":
        """Re-express a DECODED segment as raw `_val` bytes + FST outputs, the way
        Writer.Append lays them out (file/writer.go:43-56).  `encode_batch(post,
        post_off) -> (words, word_off)` is the intcomp encoder to use."""
        assert self.mode == A.II2_SEG_DECODED
        words, woff = encode_batch(self.post, self.post_off)
        return FlatSegment(self.term_bytes, self.term_off, A.II2_SEG_VAL,
                           val_bytes=words.view(np.uint8), val_off=(woff[:-1] * 4).astype(np.uint64),
                           val_size=int(woff[-1]) * 4, key=self.key)

    def with_woff32(self) -> "FlatSegment":
        """The same `_val` view with the FST outputs as 32-bit word offsets (ii2.h val_woff32)."""
        assert self.mode == A.II2_SEG_VAL and self.val_size < (1 << 34)
        return FlatSegment(self.term_bytes, self.term_off, A.II2_SEG_VAL, val_bytes=self.val_bytes,
                           val_off=None, val_size=self.val_size, key=self.key,
                           val_woff32=(self.val_off // 4).astype(np.uint32))


def views_array(segs: list[FlatSegment]):
    arr = (A.SegView * max(1, len(segs)))()
    for i, s in enumerate(segs):
        arr[i] = s.view()
    return arr


@dataclass
class MergeResult:
    terms_count: int
    term_bytes: np.ndarray
    term_off: np.ndarray
    val_off: np.ndarray
    val_bytes: np.ndarray
    val_size: int
    min_term: bytes | None
    max_term: bytes | None
    post: np.ndarray | None
    post_off: np.ndarray | None
    terms_merged: int
    postings_in: int
    postings_out: int

    @classmethod
    def from_c(cls, o: A.MergeOut, decoded: bool) -> "MergeResult":
        n = int(o.terms_count)
        term_off = A.from_ptr(o.term_off, n + 1, np.uint32)
        nb = int(term_off[-1]) if len(term_off) else 0
        return cls(
            terms_count=n,
            term_bytes=A.from_ptr(o.term_bytes, nb, np.uint8),
            term_off=term_off,
            val_off=A.from_ptr(o.val_off, n, np.uint64),
            val_bytes=A.from_ptr(o.val_bytes, int(o.val_size), np.uint8),
            val_size=int(o.val_size),
            min_term=C.string_at(o.min_term, o.min_term_len) if o.has_minmax else None,
            max_term=C.string_at(o.max_term, o.max_term_len) if o.has_minmax else None,
            post=A.from_ptr(o.post, int(o.postings_out), np.uint32) if decoded else None,
            post_off=A.from_ptr(o.post_off, n + 1, np.uint64) if decoded else None,
            terms_merged=int(o.terms_merged),
            postings_in=int(o.postings_in),
            postings_out=int(o.postings_out),
        )

    def terms(self) -> list[bytes]:
        tb = self.term_bytes.tobytes()
        return [tb[self.term_off[i]:self.term_off[i + 1]] for i in range(self.terms_count)]

    def as_dict(self) -> dict[bytes, list[int]]:
        assert self.post is not None
        return {t: self.post[int(self.post_off[i]):int(self.post_off[i + 1])].tolist()
                for i, t in enumerate(self.terms())}

    def to_segment(self, key: str = "") -> FlatSegment:
        """The merged segment as a reader would see it (`_val` + FST outputs)."""
        return FlatSegment(self.term_bytes, self.term_off, A.II2_SEG_VAL, val_bytes=self.val_bytes,
                           val_off=self.val_off, val_size=self.val_size, key=key)


@dataclass
class ReadResult:
    n_terms: int
    term_bytes: np.ndarray
    term_off: np.ndarray
    post: np.ndarray
    post_off: np.ndarray

    @classmethod
    def from_c(cls, o: A.ReadOut) -> "ReadResult":
        n = int(o.n_terms)
        term_off = A.from_ptr(o.term_off, n + 1, np.uint32)
        post_off = A.from_ptr(o.post_off, n + 1, np.uint64)
        return cls(n, A.from_ptr(o.term_bytes, int(term_off[-1]) if n else 0, np.uint8), term_off,
                   A.from_ptr(o.post, int(post_off[-1]) if n else 0, np.uint32), post_off)

    def terms(self) -> list[bytes]:
        tb = self.term_bytes.tobytes()
        return [tb[self.term_off[i]:self.term_off[i + 1]] for i in range(self.n_terms)]

    def items(self) -> list[tuple[bytes, list[int]]]:
        return [(t, self.post[int(self.post_off[i]):int(self.post_off[i + 1])].tolist())
                for i, t in enumerate(self.terms())]
