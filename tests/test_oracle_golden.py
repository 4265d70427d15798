"""Pins the CPU oracle (and the host mirror's logic) against every known-answer vector the
reference's own tests hold for the path.  CPU only."""
import numpy as np
import pytest

from inverted_index_2_b200 import _abi as A
from inverted_index_2_b200.flat import FlatSegment
from host_mirror import RemovedLists, shard_key
from scenario import (OracleBackend, check_roaring_case, load_roaring_vectors, load_vectors,
                      roaring_case_values, run_index_scenario, run_shard_scenario)

V = load_vectors()


@pytest.mark.parametrize("sc", V["shard_scenarios"], ids=lambda s: s["name"])
def test_shard_scenarios(orc, sc):
    run_shard_scenario(OracleBackend(orc), sc)


@pytest.mark.parametrize("sc", V["index_scenarios"], ids=lambda s: s["name"])
def test_index_scenarios(orc, sc):
    run_index_scenario(OracleBackend(orc), sc)


@pytest.mark.parametrize("w", V["writer"], ids=lambda s: s["name"])
def test_writer_reader_roundtrip(orc, w):
    items = [(t.encode(), v) for t, v in w["items"]]
    if w["mode"] == "direct":
        seg = FlatSegment(*FlatSegment._pack_terms([t for t, _ in items]), A.II2_SEG_DIRECT,
                          val_off=np.array([v[0] for _, v in items], dtype=np.uint64))
    else:
        seg = FlatSegment.from_items(items).to_val(orc.intcomp_encode_batch)
        # empty list -> zero bytes, the next term starts at the same offset (writer_test.go:15)
        assert seg.val_off[1] == seg.val_off[2]
    assert orc.read_range([seg]).items() == items


@pytest.mark.parametrize("b", V["bitmask"], ids=lambda s: s["name"])
def test_bitmask(orc, b):
    bm = orc.Bitmask(b["init"])
    enc = [bm.put(p) for p in b["puts"]]
    assert bm.get(enc[0] + enc[1]).tolist() == b["get_concat_first"]
    assert bm.get(enc[1]).tolist() == b["get_second_index_order"]
    assert sorted(bm.get(enc[1]).tolist()) == b["get_second_sorted"]
    assert bm.all_values().tolist() == b["all_values"]


@pytest.mark.parametrize("case", load_roaring_vectors(), ids=lambda c: c["name"])
def test_bitmask_bytes_match_roaring_format_spec(orc, case):
    """Bitmask.Put bytes (file/bitmask.go:53-59) against vectors packed straight from the public
    RoaringFormatSpec (tests/golden/make_roaring_vectors.py): array / bitmap / run containers,
    both cookies, with and without the offset header.  Spec-derived, not Go-produced: the
    container-type rules of roaring's Add() stay restated."""
    bm = orc.Bitmask(np.arange(case["dict_n"], dtype=np.uint32))
    vals = roaring_case_values(case)
    data = bm.put(vals, fast=True)
    check_roaring_case(case, data)
    assert sorted(set(bm.get(data).tolist())) == sorted(set(vals.tolist()))


def test_bitmask_out_of_bound(orc):
    big = orc.Bitmask([5, 6, 7])
    enc = big.put([7])
    small = orc.Bitmask([5])
    with pytest.raises(orc.OracleError) as e:
        small.get(enc)
    assert e.value.code == A.II2_ERR_BITMASK_OOB


def test_removed_lists():
    r = V["removed_lists"]
    rl = RemovedLists()
    rl.put(1, r["batches"][0])
    rl.put(2, r["batches"][1])
    assert rl.values().tolist() == r["values"]
    rl.sync([2, 3])
    assert rl.values().tolist() == r["after_sync_second_only"]


def test_shard_key(orc):
    for term, key in V["shard_key"]:
        assert shard_key(term.encode()) == key
        assert "%04d" % orc.shard_key(term.encode()) == key


def test_intcomp_long_list_header_chain(orc):
    """The layout facts the block-parallel long-list decoder (k3a_intcomp.cu: k_dec_tile_maps /
    _chain / _walk) is built on, checked on the oracle's own streams: the first section of a
    list of n >= 128 values is [128 * blocks, section words, first value] + blocks, a block is
    1 header + the sum of its four widths (<= 509) words, hopping from header to header from
    word 3 meets exactly `blocks` headers and ends on the section length; and the speculative
    parse — per tile, a map from every possible entry offset to (headers met, entry offset of
    the next tile), composed tile after tile — finds the same headers as the serial walk."""
    rng = np.random.default_rng(5)
    cases = [np.cumsum(rng.integers(1, 33, size=20000)).astype(np.uint32),
             rng.integers(0, 1 << 32, size=9000, dtype=np.uint64).astype(np.uint32),  # zig-zag, 32-bit widths
             np.full(8192 + 77, 5, dtype=np.uint32),                                   # one-word blocks
             np.cumsum(rng.integers(1, 1 << 13, size=8192 + 128)).astype(np.uint32)]
    tile = 1024
    for vals in cases:
        w = orc.intcomp_encode(vals)
        assert np.array_equal(orc.intcomp_decode(w), vals)
        nb, length = int(w[0]) >> 7, int(w[1])
        assert int(w[0]) == (len(vals) // 128) * 128 and int(w[2]) == int(vals[0])
        body = w[3:length].astype(np.int64)
        nxt = np.arange(len(body)) + 1 + ((body >> 24) & 0x7F) + ((body >> 16) & 0x7F) + \
            ((body >> 8) & 0x7F) + (body & 0x7F)
        # serial walk
        p, headers = 0, []
        while p < len(body):
            headers.append(p)
            assert nxt[p] - p <= 509
            p = int(nxt[p])
        assert len(headers) == nb and p == len(body)
        # speculative parse over tiles
        entry, base, found = 0, 0, []
        for t0 in range(0, len(body), tile):
            wn = min(tile, len(body) - t0)
            local = nxt[t0:t0 + wn] - t0
            maps = []
            for e in range(512):
                q, c = e, 0
                while q < wn:
                    q, c = int(local[q]), c + 1
                maps.append((c, q - wn))
            q = entry                              # the walk from the true entry
            while q < wn:
                found.append(t0 + q)
                q = int(local[q])
            c, nxt_entry = maps[entry]
            assert q - wn == nxt_entry and len(found) == base + c and nxt_entry <= 508
            entry, base = nxt_entry, base + c
        assert found == headers
