"""vellum v1 FST reader/writer (csrc/fst_v1.cpp behind ii2_fst_*) and the segment-file layer.
CPU only: the FST is host-side code.  Checked against the independent Python restatement in
oracle/vellum_ref.py (both directions) and hand-derived byte vectors; bytes are NOT pinned
against Go's vellum (module absent, no `_fst` fixtures in the reference)."""
import os
import random

import numpy as np
import pytest

from inverted_index_2_b200 import _abi as A
from inverted_index_2_b200 import files, fst
from inverted_index_2_b200.flat import FlatSegment
from oracle import vellum_ref as V

HDR = bytes([1] + [0] * 15)


def u64(x):
    return x.to_bytes(8, "little")


def test_hand_derived_bytes():
    """Byte vectors derived by hand from the v1 state encodings (fst_v1.cpp header)."""
    # no keys: a non-final root without transitions = pack 0x00, count byte 0x00 (a count of 0
    # does not fit the header's 1..63), header 0x00
    assert fst.fst_build_items([]) == HDR + bytes([0x00, 0x00, 0x00]) + u64(0) + u64(18)
    # {"t": 0}: root has one transition on a common input ('t' = code 1) to the implicit final
    # state (address 0): delta 0 in one byte, pack 0x10, header 0x80 | 1
    assert fst.fst_build_items([(b"t", 0)]) == HDR + bytes([0x00, 0x10, 0x81]) + u64(1) + u64(18)
    # {"~": 5}: uncommon input byte stored below the header; output 5 in one byte
    assert fst.fst_build_items([(b"~", 5)]) == HDR + bytes([0x05, 0x00, 0x11, 0x7E, 0x80]) + u64(1) + u64(20)
    # only the empty key with value 0: the root IS the implicit final state
    assert fst.fst_build_items([(b"", 0)]) == HDR + u64(1) + u64(0)
    # {"ab": 1, "ac": 2}: state after 'a' has two transitions (both to address 0) with outputs
    # 0 and 1 after the common prefix took min(1, 2) = 1; root -> it is "next state" form
    b = fst.fst_build_items([(b"ab", 1), (b"ac", 2)])
    mid = bytes([0x01, 0x00,   # outputs of 'c', 'b' (reversed)
                 0x00, 0x00,   # deltas of 'c', 'b'
                 ord("c"), ord("b"), 0x11, 0x02])
    root = bytes([0x01, 0x01, 0x11, 0x80 | 5])  # out 1, delta 24 - 23 = 1, pack, 'a' = code 5
    assert b == HDR + mid + root + u64(2) + u64(16 + len(mid) + len(root) - 1)


def random_items(rng, n, alphabet, maxlen, vmax):
    keys = set()
    while len(keys) < n:
        keys.add(bytes(rng.choice(alphabet) for _ in range(rng.randint(0, maxlen))))
    return [(k, rng.randint(0, vmax)) for k in sorted(keys)]


@pytest.mark.parametrize("seed,n,alphabet,maxlen,vmax", [
    (1, 300, b"ab", 10, 5), (2, 2000, b"abcdefghijklmnopqrstuvwxyz", 8, 1 << 20),
    (3, 500, bytes(range(256)), 4, (1 << 64) - 1), (4, 1000, b"te/oasr~\x00\xff", 12, 1 << 33),
])
def test_roundtrip_against_python_restatement(seed, n, alphabet, maxlen, vmax):
    rng = random.Random(seed)
    items = random_items(rng, n, alphabet, maxlen, vmax)
    data = fst.fst_build_items(items)
    assert V.decode(data) == items            # C++ writer -> Python reader
    assert fst.fst_items(data) == items       # C++ writer -> C++ reader
    assert fst.fst_items(V.encode_trie(items)) == items  # Python writer -> C++ reader
    assert fst.fst_len(data) == len(items)
    # suffix sharing: the minimised automaton is smaller than the plain trie
    assert len(data) <= len(V.encode_trie(items))
    for k, v in rng.sample(items, 50):
        assert fst.fst_get(data, k) == v
    assert fst.fst_get(data, b"\x01no such key\x02") is None
    # range reads = Iterator(min, nil) + the reader's inclusive max
    for _ in range(30):
        lo = rng.choice(items)[0] if rng.random() < 0.5 else bytes(rng.choice(alphabet) for _ in range(3))
        hi = rng.choice(items)[0] if rng.random() < 0.5 else bytes(rng.choice(alphabet) for _ in range(3))
        lo = None if rng.random() < 0.2 else lo
        hi = None if rng.random() < 0.2 else hi
        exp = [(k, v) for k, v in items if (lo is None or k >= lo) and (hi is None or k <= hi)]
        assert fst.fst_items(data, lo, hi) == exp


def test_wide_state_and_many_keys():
    # 256 transitions out of the root (count stored as 1), all to the implicit final state
    items = [(bytes([b]), b * 3) for b in range(256)]
    data = fst.fst_build_items(items)
    assert V.decode(data) == items and fst.fst_items(data) == items
    # 100 transitions: count byte below the header
    items = [(bytes([b]), 7) for b in range(100)]
    assert fst.fst_items(fst.fst_build_items(items)) == items
    # 200k direct-mode style keys with one shared value (Shard.Put, shard.go:33-67)
    rng = random.Random(9)
    keys = sorted({bytes(rng.choice(b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ")
                         for _ in range(rng.randint(10, 19))) for _ in range(200000)})
    off = np.zeros(len(keys) + 1, dtype=np.uint32)
    off[1:] = np.cumsum([len(k) for k in keys])
    tb = np.frombuffer(b"".join(keys), dtype=np.uint8)
    data = fst.fst_build(tb, off, np.full(len(keys), 42, dtype=np.uint64))
    gtb, goff, gval, glen = fst.fst_read(data)
    assert glen == len(keys) and np.array_equal(goff, off) and np.array_equal(gtb, tb)
    assert (gval == 42).all()


def test_rejects_bad_input():
    with pytest.raises(fst.FstError) as e:
        fst.fst_build_items([(b"b", 1), (b"a", 2)])
    assert e.value.code == A.II2_ERR_INVALID
    with pytest.raises(fst.FstError):
        fst.fst_build_items([(b"a", 1), (b"a", 2)])
    with pytest.raises(fst.FstError) as e:
        fst.fst_read(b"short")
    assert e.value.code == A.II2_ERR_CORRUPT
    good = fst.fst_build_items([(b"abc", 1), (b"abd", 2)])
    bad = bytes([2]) + good[1:]
    with pytest.raises(fst.FstError) as e:
        fst.fst_read(bad)
    assert e.value.code == A.II2_ERR_UNSUPPORTED
    trunc = good[:-8] + u64(len(good) + 100)  # root address outside the file
    with pytest.raises(fst.FstError):
        fst.fst_read(trunc)


def test_segment_files_roundtrip(tmp_path, orc):
    """file/writer_test.go:13-45 and :52-83 through real files: `<key>_fst` + `<key>_val`
    (full mode, unsorted and empty lists preserved) and FST-only direct mode."""
    d = str(tmp_path)
    items = [(b"term1", [10, 500, 300]), (b"term2", []), (b"term3", [66, 5513])]
    seg = FlatSegment.from_items(items).to_val(orc.intcomp_encode_batch)
    files.write_segment(d, "100", seg)
    assert sorted(os.listdir(d)) == ["100_fst", "100_val"]
    back = files.open_segment(d, "100")
    assert back.mode == A.II2_SEG_VAL and back.terms() == [t for t, _ in items]
    assert back.val_off[1] == back.val_off[2]  # empty list: zero bytes (writer_test.go:15)
    assert orc.read_range([back]).items() == items
    # ranged opens keep the run-length rule (next offset or file size)
    assert orc.read_range([files.open_segment(d, "100", b"term2", b"term2")]).items() == [items[1]]
    assert orc.read_range([files.open_segment(d, "100", b"term3", None)]).items() == [items[2]]
    assert orc.read_range([files.open_segment(d, "100", None, b"term1")]).items() == [items[0]]
    assert files.open_segment(d, "100", b"u", None) is None
    assert files.open_segment(d, "100", None, b"a") is None
    direct = FlatSegment(*FlatSegment._pack_terms([b"term1", b"term2"]), A.II2_SEG_DIRECT,
                         val_off=np.array([10, 11], dtype=np.uint64))
    files.write_segment(d, "101", direct)
    assert not os.path.exists(os.path.join(d, "101_val"))
    back = files.open_segment(d, "101")
    assert back.mode == A.II2_SEG_DIRECT
    assert orc.read_range([back]).items() == [(b"term1", [10]), (b"term2", [11])]
    assert files.list_segments(d) == ["100", "101"]
    files.remove_segment(d, "100")
    assert files.list_segments(d) == ["101"]


def test_index_directory_reopen(tmp_path, orc):
    """inverted_index_test.go:140-194: the index written through segment files (one directory
    per shard, `<key>_fst` [+ `<key>_val`]) answers the same after it is reopened; merged
    segments replace their sources on disk (shard.go:232-242)."""
    from host_mirror import InvertedIndex
    from scenario import OracleBackend
    d = str(tmp_path)
    idx = InvertedIndex(OracleBackend(orc), basedir=d)
    idx.put([b"ab1", b"ab2"], 1)
    idx.put([b"ab2", b"cd1"], 2)
    exp = [(b"ab1", [1]), (b"ab2", [1, 2]), (b"cd1", [2])]
    assert list(idx.read(None, None)) == exp and len(idx.shards) == 2
    again = InvertedIndex(OracleBackend(orc), basedir=d)
    assert len(again.shards) == 2 and list(again.read(None, None)) == exp
    assert again.merge(2, 10) == 2  # the two direct segments of shard "ab" become one full one
    shard_dirs = sorted(os.listdir(d))
    ab = sorted(os.listdir(os.path.join(d, shard_dirs[0])))
    assert len(ab) == 2 and ab[0].endswith("_fst") and ab[1].endswith("_val")
    third = InvertedIndex(OracleBackend(orc), basedir=d)
    assert list(third.read(None, None)) == exp
    assert third.prefix_search([b"ab", b"zz"]) == {b"ab": [1, 2]}


# ---------------------------------------------------------------- removed.list (gob)
def test_gob_removed_list_roundtrip_and_known_bytes():
    """removed_list_test.go:9-18 through the gob stream; the byte layout is checked against the
    encodings the encoding/gob documentation fixes (uint / int forms, message framing, the
    singleton zero byte), not against a file written by Go."""
    lists = {1: np.array([1, 5, 10], dtype=np.uint32), 2: np.array([2, 20, 30], dtype=np.uint32)}
    data = fst.removed_list_encode(lists)
    back = fst.removed_list_decode(data)
    assert sorted(back) == [1, 2] and all(np.array_equal(back[k], lists[k]) for k in lists)
    # message 1: type -66 = map{Id 66, Key int(2), Elem 65}; message 2: type -65 = slice{Id 65,
    # Elem uint(3)}; message 3: value of type 66: 0 (singleton), 2 entries
    m1 = bytes([0xFF, 0x83, 0x04, 0x01, 0x02, 0xFF, 0x84, 0x00, 0x01, 0x04, 0x01, 0xFF, 0x82, 0x00, 0x00])
    m2 = bytes([0xFF, 0x81, 0x02, 0x01, 0x02, 0xFF, 0x82, 0x00, 0x01, 0x06, 0x00, 0x00])
    m3 = bytes([0xFF, 0x84, 0x00, 0x02, 0x02, 0x03, 0x01, 0x05, 0x0A, 0x04, 0x03, 0x02, 0x14, 0x1E])
    assert data == bytes([len(m1)]) + m1 + bytes([len(m2)]) + m2 + bytes([len(m3)]) + m3
    # wide values: unix-nano timestamps (9 bytes on the wire), negative keys, values >= 2^31
    big = {1_700_000_000_123_456_789: np.array([0, 127, 128, 0xFFFFFFFF], dtype=np.uint32),
           -5: np.zeros(0, dtype=np.uint32), 0: np.arange(300, dtype=np.uint32)}
    back = fst.removed_list_decode(fst.removed_list_encode(big))
    assert sorted(back) == sorted(big) and all(np.array_equal(back[k], big[k]) for k in big)
    assert fst.removed_list_decode(fst.removed_list_encode({})) == {}


def test_gob_decoder_accepts_named_types_and_other_ids():
    """A stream as another gob encoder may write it: type names present, different type ids,
    the slice described before the map."""
    def msg(body):
        assert len(body) < 128
        return bytes([len(body)]) + bytes(body)
    name_s, name_m = b"[]uint32", b"map[int64][]uint32"
    # ids 70 (slice) and 71 (map): int(-70) = 139, int(70) = 140, int(-71) = 141, int(71) = 142
    slice_def = [0xFF, 139, 0x02, 0x01, 0x01, len(name_s), *name_s, 0x01, 0xFF, 140, 0x00,
                 0x01, 0x06, 0x00, 0x00]
    map_def = [0xFF, 141, 0x04, 0x01, 0x01, len(name_m), *name_m, 0x01, 0xFF, 142, 0x00,
               0x01, 0x04, 0x01, 0xFF, 140, 0x00, 0x00]
    value = [0xFF, 142, 0x00, 0x01, 0x0E, 0x02, 0x09, 0xFE, 0x01, 0x00]  # {7: [9, 256]}
    got = fst.removed_list_decode(msg(slice_def) + msg(map_def) + msg(value))
    assert list(got) == [7] and got[7].tolist() == [9, 256]
    with pytest.raises(fst.FstError):
        fst.removed_list_decode(msg(slice_def) + msg(map_def) + msg(value)[:-2])
    with pytest.raises(fst.FstError):
        fst.removed_list_decode(msg(value))  # value of an undescribed type


def test_removed_list_persists_with_the_shard(tmp_path, orc):
    from host_mirror import InvertedIndex
    from scenario import OracleBackend
    d = str(tmp_path)
    idx = InvertedIndex(OracleBackend(orc), basedir=d)
    idx.put([b"aaaa", b"bbbb"], 1)
    idx.put([b"aaaa", b"bbbb"], 1)
    idx.put([b"aaaa"], 2)
    idx.put_removed([1])
    shard_dir = os.path.join(d, sorted(os.listdir(d))[0])
    assert "removed.list" in os.listdir(shard_dir)
    again = InvertedIndex(OracleBackend(orc), basedir=d)  # removed list reloaded from disk
    assert again.shards[0].removed_list.values().tolist() == [1]
    assert again.merge(2, 3, 2) > 0
    assert list(again.read(None, None)) == [(b"aaaa", [2])]  # inverted_index_test.go:59-82
