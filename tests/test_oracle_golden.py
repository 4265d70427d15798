"""Pins the CPU oracle (and the host mirror's logic) against every known-answer vector the
reference's own tests hold for the path.  CPU only."""
import numpy as np
import pytest

from inverted_index_2_b200 import _abi as A
from inverted_index_2_b200.flat import FlatSegment
from inverted_index_2_b200.host import RemovedLists, shard_key
from scenario import OracleBackend, load_vectors, run_index_scenario, run_shard_scenario

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
