"""Randomised differential test: the CUDA path behind the C-ABI against the CPU oracle on
generated shard contents — skewed term alphabets with long shared prefixes, empty and very
uneven segments, empty lists, unsorted single-source lists (survey Q4), duplicates across
segments, values up to 2^32-1, removed lists with and without a bitmap, open / closed / absent
range bounds, point reads — on both bucket pipelines.  Deterministic (fixed seeds): a failure reproduces."""
import numpy as np
import pytest

from inverted_index_2_b200.flat import FlatSegment

pytestmark = pytest.mark.gpu


def _terms(rng, n, style):
    if style == "prefix":          # long common prefixes, lengths 1 .. 40
        base = [b"", b"a", b"ab", b"abc" * 5, b"zz" * 12]
        out = {bytes(rng.choice(base)) + bytes(rng.integers(97, 100, size=int(rng.integers(0, 6))).tolist())
               for _ in range(n)}
    elif style == "binary":        # every byte value, including 0x00 and 0xff, and the empty term
        out = {bytes(rng.integers(0, 256, size=int(rng.integers(0, 5))).tolist()) for _ in range(n)}
    else:                          # the reference's generator shape (shard_test.go:258-266)
        out = {bytes(rng.choice(list(b"abcXYZ"), size=int(rng.integers(2, 12))).tolist()) for _ in range(n)}
    return sorted(out)


def _segments(rng, vocab, nseg, hi_bits):
    segs = []
    for s in range(nseg):
        frac = rng.choice([0.0, 0.05, 0.5, 1.0])
        items = []
        for t in vocab:
            if rng.random() >= frac:
                continue
            n = int(rng.choice([0, 1, 2, 3, 7, 40, 130, 300], p=[.05, .3, .2, .2, .1, .1, .03, .02]))
            vals = rng.integers(0, 1 << hi_bits, size=n, dtype=np.int64)
            if rng.random() < 0.7:
                vals = np.unique(vals)      # the usual case: sorted unique lists
            items.append((t, vals.tolist()))  # else: unsorted, with duplicates (writer_test.go:14)
        segs.append(FlatSegment.from_items(items))
    return segs


import os

# II2_FUZZ_SEEDS=<n> widens the run (a one-off before a release; the default stays short)
@pytest.mark.parametrize("seed", range(int(os.environ.get("II2_FUZZ_SEEDS", "24"))))
@pytest.mark.parametrize("path", ["0", "1"])
def test_random_shard_contents(engine, orc, monkeypatch, seed, path):
    monkeypatch.setenv("II2_FUSED", path)
    rng = np.random.default_rng(1000 + seed)
    style = ["prefix", "binary", "alpha"][seed % 3]
    hi_bits = [8, 16, 32][(seed // 3) % 3]
    vocab = _terms(rng, int(rng.integers(1, 400)), style)
    segs = _segments(rng, vocab, int(rng.integers(1, 20)), hi_bits)
    nrem = int(rng.choice([0, 3, 100, 5000]))
    removed = np.unique(rng.integers(0, 1 << hi_bits, size=nrem, dtype=np.int64)).astype(np.uint32) if nrem else None
    got, exp = engine.merge(segs, removed, decoded=True), orc.merge(segs, removed, decoded=True)
    for f in ("terms_count", "val_size", "min_term", "max_term", "terms_merged", "postings_in", "postings_out"):
        assert getattr(got, f) == getattr(exp, f), f
    for f in ("term_bytes", "term_off", "val_off", "val_bytes", "post", "post_off"):
        assert np.array_equal(getattr(got, f), getattr(exp, f)), f
    for _ in range(3):
        lo = None if rng.random() < 0.3 else bytes(rng.choice(vocab)) if rng.random() < 0.6 else b"ab~"
        hi = None if rng.random() < 0.3 else bytes(rng.choice(vocab)) if rng.random() < 0.6 else b"b"
        r, e = engine.read_range(segs, lo, hi), orc.read_range(segs, lo, hi)
        assert r.n_terms == e.n_terms
        for f in ("term_bytes", "term_off", "post", "post_off"):
            assert np.array_equal(getattr(r, f), getattr(e, f)), (f, lo, hi)
    # point reads (min == max): a term of the vocabulary and one that is absent, with the filter
    for t in (bytes(rng.choice(vocab)), bytes(rng.choice(vocab)) + b"\x01"):
        r, e = engine.read_range(segs, t, t, removed=removed), orc.read_range(segs, t, t, removed=removed)
        assert r.n_terms == e.n_terms
        for f in ("term_bytes", "term_off", "post", "post_off"):
            assert np.array_equal(getattr(r, f), getattr(e, f)), (f, t)


@pytest.mark.parametrize("seed", range(int(os.environ.get("II2_FUZZ_HEAVY_SEEDS", "8"))))
def test_random_heavy_terms(engine, orc, seed):
    """Few terms with long lists over many segments: the warp- and CTA-per-term unions (257 ...
    4096 values, both gathers, both CTA shapes), the multi-CTA path beyond, next to light terms in
    the same buckets; merge (both outputs) and a range read with the filter."""
    rng = np.random.default_rng(7000 + seed)
    hi_bits = [12, 20, 32][seed % 3]
    vocab = _terms(rng, int(rng.integers(2, 30)), "alpha")
    nseg = int(rng.integers(2, 48))
    sizes = [0, 1, 5, 40, 130, 300, 900, 2500]
    weights = np.array([.05, .15, .15, .2, .15, .15, .1, .05])
    segs = []
    for s in range(nseg):
        frac = rng.choice([0.2, 0.6, 1.0])
        items = []
        for t in vocab:
            if rng.random() >= frac:
                continue
            n = int(rng.choice(sizes, p=weights))
            vals = rng.integers(0, 1 << hi_bits, size=n, dtype=np.int64)
            if rng.random() < 0.8:
                vals = np.unique(vals)
            items.append((t, vals.tolist()))
        segs.append(FlatSegment.from_items(items))
    nrem = int(rng.choice([0, 50, 3000]))
    removed = np.unique(rng.integers(0, 1 << hi_bits, size=nrem, dtype=np.int64)).astype(np.uint32) if nrem else None
    got, exp = engine.merge(segs, removed, decoded=True), orc.merge(segs, removed, decoded=True)
    for f in ("terms_count", "val_size", "terms_merged", "postings_in", "postings_out"):
        assert getattr(got, f) == getattr(exp, f), f
    for f in ("term_bytes", "term_off", "val_off", "val_bytes", "post", "post_off"):
        assert np.array_equal(getattr(got, f), getattr(exp, f)), f
    lo, hi = vocab[len(vocab) // 4], vocab[-1]
    r, e = engine.read_range(segs, lo, hi, removed=removed), orc.read_range(segs, lo, hi, removed=removed)
    assert r.n_terms == e.n_terms
    for f in ("term_bytes", "term_off", "post", "post_off"):
        assert np.array_equal(getattr(r, f), getattr(e, f)), f
