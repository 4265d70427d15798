"""Deterministic synthetic workloads for the parity tests and bench.py.

`terms.1m.txt` (BASELINE.json) is a missing blob of the reference, so the term set is the
reference's own generator shape: random strings over a-zA-Z with length uniform in [10,20)
(shard_test.go:258-266, used as randomString(10,20) at inverted_index_test.go:98-100), made
unique and sorted by bytes.Compare.  Everything is numpy + PCG64 with fixed seeds.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi as A
from .flat import FlatSegment

ALPHABET = np.frombuffer(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ", dtype=np.uint8)


def make_terms(n: int, seed: int = 0x1EE7, lo: int = 10, hi: int = 20):
    """n unique terms, sorted.  Returns (term_bytes u8, term_off u32[n+1])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    width = hi  # fixed-width, NUL padded: numpy 'S' ordering == bytes.Compare for NUL-free terms
    out = np.zeros(0, dtype=f"S{width}")
    while len(out) < n:
        m = int((n - len(out)) * 1.05) + 16
        mat = ALPHABET[rng.integers(0, len(ALPHABET), size=(m, width))]
        lens = rng.integers(lo, hi, size=m)
        mat[np.arange(width)[None, :] >= lens[:, None]] = 0
        out = np.unique(np.concatenate([out, mat.view(f"S{width}").ravel()]))
    if len(out) > n:  # drop random extras, keep order
        keep = np.sort(rng.choice(len(out), size=n, replace=False))
        out = out[keep]
    mat = out.view(np.uint8).reshape(n, width)
    lens = (mat != 0).sum(axis=1).astype(np.uint32)
    off = np.zeros(n + 1, dtype=np.uint32)
    np.cumsum(lens, out=off[1:])
    tb = mat[mat != 0].copy()  # row-major: terms stay contiguous and in order
    return tb, off


def gather_terms(tb: np.ndarray, off: np.ndarray, idx: np.ndarray):
    """Term bytes/offsets of the sub-dictionary idx (ascending indexes)."""
    lens = (off[idx + 1] - off[idx]).astype(np.int64)
    noff = np.zeros(len(idx) + 1, dtype=np.uint32)
    np.cumsum(lens, out=noff[1:])
    total = int(noff[-1])
    # byte j of output term i comes from off[idx[i]] + j
    src = np.repeat(off[idx].astype(np.int64) - noff[:-1].astype(np.int64), lens) + np.arange(total)
    return tb[src], noff


@dataclass
class Workload:
    term_bytes: np.ndarray
    term_off: np.ndarray
    segments: list[FlatSegment]
    seg_term_ids: list[np.ndarray]  # global term id of every term of every segment
    removed: np.ndarray  # sorted uint32
    universe: int
    postings_in: int
    term_instances: int

    def expected_union(self, removed: np.ndarray | None = None):
        """Independent numpy answer for the merge of ALL segments: sorted unique
        (term id, value) pairs, removed values dropped, as (term_ids, post, post_off)
        over surviving terms.  Valid because every generated list is sorted-unique, so
        the single-source pass-through (survey Q4) is invisible."""
        keys = []
        for seg, ids in zip(self.segments, self.seg_term_ids):
            lens = np.diff(seg.post_off).astype(np.int64)
            keys.append((np.repeat(ids.astype(np.uint64), lens) << np.uint64(32)) |
                        seg.post.astype(np.uint64))
        k = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.uint64)
        vals = (k & np.uint64(0xFFFFFFFF)).astype(np.uint32)
        if removed is not None and len(removed):
            keep = ~np.isin(vals, removed)
            k, vals = k[keep], vals[keep]
        tid = (k >> np.uint64(32)).astype(np.int64)
        terms, counts = np.unique(tid, return_counts=True)
        post_off = np.zeros(len(terms) + 1, dtype=np.uint64)
        np.cumsum(counts, out=post_off[1:])
        return terms, vals, post_off


def make_workload(n_terms: int, n_segments: int, postings: int, *, seed: int = 0xC2,
                  presence: float = 0.5, universe: int = 1 << 24, removed_frac: float = 0.05,
                  removed_seed: int = 0xDE1, terms_seed: int = 0x1EE7, max_len: int = 64,
                  terms=None) -> Workload:
    """SURVEY.md §8(d) C2/C3 shape: segment s holds term t with probability `presence`;
    list length = geometric (>= 1) with the mean that makes the total ~= `postings`; values
    are sorted unique draws from [0, universe)."""
    tb, off = terms if terms is not None else make_terms(n_terms, terms_seed)
    n_terms = len(off) - 1
    rng = np.random.Generator(np.random.PCG64(seed))
    mean_len = max(1.0, postings / max(1.0, n_terms * n_segments * presence))
    p = min(1.0, 1.0 / mean_len)
    gap_hi = max(2, (universe // 2) // max_len)
    segs, ids_all = [], []
    total_post = total_inst = 0
    for s in range(n_segments):
        ids = np.nonzero(rng.random(n_terms) < presence)[0]
        lens = np.minimum(rng.geometric(p, size=len(ids)), max_len).astype(np.int64)
        poff = np.zeros(len(ids) + 1, dtype=np.uint64)
        np.cumsum(lens, out=poff[1:])
        n = int(poff[-1])
        gaps = rng.integers(1, gap_hi + 1, size=n, dtype=np.int64)
        first = rng.integers(0, universe // 2, size=len(ids), dtype=np.int64)
        c = np.cumsum(gaps)
        starts = poff[:-1].astype(np.int64)
        base = np.repeat(c[starts] - first, lens) if n else np.zeros(0, dtype=np.int64)
        post = (c - base).astype(np.uint32)  # first value of list i == first[i], then + gaps
        stb, stoff = gather_terms(tb, off, ids)
        segs.append(FlatSegment(stb, stoff, A.II2_SEG_DECODED, post=post, post_off=poff,
                                key=str(s)))
        ids_all.append(ids)
        total_post += n
        total_inst += len(ids)
    rrng = np.random.Generator(np.random.PCG64(removed_seed))
    nrem = int(universe * removed_frac)
    removed = np.sort(rrng.choice(universe, size=nrem, replace=False).astype(np.uint32)) \
        if nrem else np.zeros(0, dtype=np.uint32)
    return Workload(tb, off, segs, ids_all, removed, universe, total_post, total_inst)


def term_at(tb: np.ndarray, off: np.ndarray, i: int) -> bytes:
    return tb[off[i]:off[i + 1]].tobytes()


def algorithmic_bytes(n_in: int, n_out: int, t_in: int, t_in_bytes: int, n_segs: int, t_out: int,
                      t_out_bytes: int, n_removed: int) -> int:
    """every datum of the decoded domain touched once."""
    return (4 * n_in + 4 * n_out + t_in_bytes + 4 * (t_in + n_segs) + t_out_bytes +
            4 * (t_out + 1) + 8 * (t_in + n_segs) + 8 * (t_out + 1) + 4 * n_removed)
