"""Full-size parity (BASELINE.json configs at their stated sizes): the CPU oracle needs minutes
at these sizes, so the CUDA path is checked against an INDEPENDENT numpy answer (np.unique over
(term id, value) pairs — no oracle code, no library code) and through size-independent
properties of the domain: strictly ascending terms, sorted-unique lists, no removed value
survives, `_val` decodes back to the postings, merging is associative over pass structure and
idempotent on its own result, bitmask / intcomp round trips.  Integer / byte work: bit-exact.

  C2  compaction of 64 segments, 1 M terms, 100 M postings, 5 % of the id universe removed
  C3  term-range read union across 256 segments with the 5 % filter (read, then
      the Merge-style filter)
  C4  bitmask / intcomp at 16 M values (the top of the sweep)
"""
import numpy as np
import pytest

from inverted_index_2_b200 import synth

pytestmark = pytest.mark.gpu


def _expected(w):
    terms, vals, poff = w.expected_union(w.removed)
    etb, eoff = synth.gather_terms(w.term_bytes, w.term_off, terms)
    return etb, eoff, vals, poff


def _check_list_properties(post, post_off, removed):
    """Every list strictly ascending (slices.Sort + slices.Compact, file/types.go:14-22) and
    free of removed values (shard.go:181-190)."""
    starts = np.zeros(len(post), dtype=bool)
    starts[post_off[:-1][post_off[:-1] < len(post)].astype(np.int64)] = True
    d = np.diff(post.astype(np.int64))
    assert bool(np.all((d > 0) | starts[1:]))
    assert not bool(np.isin(post, removed).any())


def _terms_strictly_ascending(tb, off):
    """bytes.Compare order on the first 8 bytes + length tiebreak is enough for a strictness
    check of random terms; equal 8-byte prefixes fall back to a Python compare."""
    n = len(off) - 1
    key = np.zeros(n, dtype=np.uint64)
    lens = np.diff(off.astype(np.int64))
    for b in range(8):
        has = lens > b
        byte = np.zeros(n, dtype=np.uint64)
        byte[has] = tb[(off[:-1].astype(np.int64) + b)[has]]
        key = (key << np.uint64(8)) | byte
    lt = key[:-1] < key[1:]
    ties = np.nonzero(~lt)[0]
    for i in ties.tolist():
        a = tb[off[i]:off[i + 1]].tobytes()
        b2 = tb[off[i + 1]:off[i + 2]].tobytes()
        assert a < b2, (i, a, b2)
    return True


def test_c2_compaction_full_size(engine):
    w = synth.make_workload(1_000_000, 64, 100_000_000, seed=0xC2, removed_frac=0.05)
    dsegs = [engine.upload(s) for s in w.segments]
    drem = engine.upload_removed(w.removed)
    one_dev = engine.merge_dev(dsegs, drem, encode=True, decoded=True)
    one = one_dev.download_merge(decoded=True)
    # (1) the independent answer
    etb, eoff, vals, poff = _expected(w)
    assert one.terms_count == len(eoff) - 1
    assert np.array_equal(one.term_off, eoff) and np.array_equal(one.term_bytes, etb)
    assert np.array_equal(one.post_off, poff) and np.array_equal(one.post, vals)
    assert one.postings_in == w.postings_in and one.postings_out == len(vals)
    assert one.terms_merged == len(w.term_off) - 1
    # (2) properties
    assert _terms_strictly_ascending(one.term_bytes, one.term_off)
    _check_list_properties(one.post, one.post_off, w.removed)
    # (3) `_val` is one intcomp stream per term at the FST offsets (file/writer.go:43-56)
    assert one.val_size % 4 == 0 and int(one.val_off[0]) == 0
    woff = np.concatenate([one.val_off // 4, [one.val_size // 4]]).astype(np.uint64)
    dec, doff = engine.intcomp_decode_batch(one.val_bytes.view(np.uint32), woff)
    assert np.array_equal(doff, one.post_off) and np.array_equal(dec, one.post)
    # (4) pass structure is invisible: (32 + 32 segments) then the two results == one pass
    r1 = engine.merge_dev(dsegs[:32], None, encode=False).to_segment()
    r2 = engine.merge_dev(dsegs[32:], None, encode=False).to_segment()
    two = engine.merge_dev([r1, r2], drem, encode=True, decoded=True).download_merge(decoded=True)
    assert np.array_equal(two.val_bytes, one.val_bytes) and np.array_equal(two.val_off, one.val_off)
    assert np.array_equal(two.term_bytes, one.term_bytes) and np.array_equal(two.post, one.post)
    # (5) idempotence: the merged segment merged again (alone, and with the same filter)
    again = engine.merge_dev([one_dev.to_segment()], drem, encode=True, decoded=True) \
        .download_merge(decoded=True)
    assert np.array_equal(again.val_bytes, one.val_bytes)
    assert np.array_equal(again.term_bytes, one.term_bytes) and np.array_equal(again.post, one.post)
    # (6) the host-buffer entry point (pipelined by term range) returns the same bytes
    host = engine.merge(w.segments, w.removed, decoded=False)
    assert np.array_equal(host.val_bytes, one.val_bytes) and np.array_equal(host.val_off, one.val_off)
    assert np.array_equal(host.term_bytes, one.term_bytes) and np.array_equal(host.term_off, one.term_off)


def test_c3_range_read_256_segments_full_size(engine):
    w = synth.make_workload(1_000_000, 256, 100_000_000, seed=0xC3, removed_frac=0.05)
    dsegs = [engine.upload(s) for s in w.segments]
    drem = engine.upload_removed(w.removed)
    rd = engine.read_range_dev(dsegs, None, None, drem).download_read()
    etb, eoff, vals, poff = _expected(w)
    # a read keeps a term whose list the filter emptied (only Merge drops it, shard.go:192-194)
    lens = np.diff(rd.post_off.astype(np.int64))
    assert np.array_equal(rd.post, vals)
    keep = lens > 0
    assert int(keep.sum()) == len(eoff) - 1
    assert np.array_equal(rd.post_off[:-1][keep], poff[:-1])
    tl = np.diff(rd.term_off.astype(np.int64))
    assert np.array_equal(tl[keep], np.diff(eoff.astype(np.int64)))
    if bool(keep.all()):
        assert np.array_equal(rd.term_bytes, etb) and np.array_equal(rd.term_off, eoff)
    _check_list_properties(rd.post, rd.post_off, w.removed)
    # a 1 % range equals the same slice of the full read (inclusive bounds, file/reader.go:54-58)
    n = len(w.term_off) - 1
    a, b = n // 2, n // 2 + n // 100
    lo, hi = synth.term_at(w.term_bytes, w.term_off, a), synth.term_at(w.term_bytes, w.term_off, b)
    part = engine.read_range_dev(dsegs, lo, hi, drem).download_read()
    assert rd.n_terms == n and part.n_terms == b - a + 1  # every term is in some segment
    p0, p1 = int(rd.post_off[a]), int(rd.post_off[b + 1])
    assert np.array_equal(part.post, rd.post[p0:p1])
    assert np.array_equal(part.post_off, rd.post_off[a:b + 2] - rd.post_off[a])
    assert np.array_equal(part.term_bytes, rd.term_bytes[int(rd.term_off[a]):int(rd.term_off[b + 1])])


def test_c4_codecs_at_16m_values(engine):
    L = 1 << 24
    rng = np.random.default_rng(0xB17)
    # bitmask: dictionary = sorted universe of 2L ids, values = a random half (bitmask_test.go:15-21)
    universe = np.arange(2 * L, dtype=np.uint32) * 3 + 1
    vals = universe[rng.permutation(2 * L)[:L]]
    bm = engine.bitmask(universe)
    enc = bm.put(vals)
    got = bm.get(enc)  # ascending INDEX order == ascending value order for a sorted dictionary
    assert np.array_equal(got, np.sort(vals))
    assert len(enc) <= 16 + (2 * L // 65536) * (8192 + 12)  # bitmap containers + header
    assert bm.put(vals[::-1].copy()) == enc  # a set: order and repetition of the input are invisible
    # intcomp: one 16 M list, and 16 M values cut into ragged lists (sorted and unsorted)
    one = np.cumsum(rng.integers(1, 128, size=L, dtype=np.int64)).astype(np.uint32)  # sorted unique
    words, woff = engine.intcomp_encode_batch(one, np.array([0, L], dtype=np.uint64))
    dec, doff = engine.intcomp_decode_batch(words, woff)
    assert np.array_equal(dec, one) and doff.tolist() == [0, L]
    assert len(words) < L  # gaps below 128 pack far below 32 bits per value
    cuts = np.unique(np.concatenate([[0, L], rng.integers(0, L, size=200_000)])).astype(np.uint64)
    unsorted = rng.integers(0, 1 << 32, size=L, dtype=np.uint64).astype(np.uint32)
    for data in (one, unsorted):
        words, woff = engine.intcomp_encode_batch(data, cuts)
        dec, doff = engine.intcomp_decode_batch(words, woff)
        assert np.array_equal(doff, cuts) and np.array_equal(dec, data)
