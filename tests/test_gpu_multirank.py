"""Two ranks on two GPUs: the cross-shard exchange behind the C-ABI (ii2_comm_init,
ii2_read_gather, ii2_prefix_gather — NCCL over NVLink) against the single-process answer.
Needs >= 2 GPUs (`gpurun --gpus 2`); skipped otherwise."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _workload():
    from inverted_index_2_b200 import synth
    return synth.make_workload(30000, 8, 400000, universe=1 << 18, seed=77)


def _rank_segments(w, rank, world):
    """Term-range partition by shard key: contiguous ranges of the sorted term list."""
    from inverted_index_2_b200 import sharded
    from inverted_index_2_b200.flat import FlatSegment
    keys = sharded.shard_keys_of_sorted(w.term_bytes, w.term_off)
    bounds = sharded.partition_shard_keys(np.bincount(keys, minlength=1024).astype(float), world)
    lo, hi = np.searchsorted(keys, [bounds[rank], bounds[rank + 1]])
    out = []
    for seg, ids in zip(w.segments, w.seg_term_ids):
        a, b = np.searchsorted(ids, [lo, hi])
        toff, poff = seg.term_off[a:b + 1], seg.post_off[a:b + 1]
        out.append(FlatSegment(seg.term_bytes[int(toff[0]):int(toff[-1])].copy(),
                               (toff - toff[0]).astype(np.uint32), seg.mode,
                               post=seg.post[int(poff[0]):int(poff[-1])].copy(),
                               post_off=(poff - poff[0]).astype(np.uint64)))
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from inverted_index_2_b200 import sharded, synth
    from inverted_index_2_b200.engine import Engine
    eng = Engine(rank)
    w = _workload()
    dsegs = [eng.upload(s) for s in _rank_segments(w, rank, world)]
    sharded.comm_init_from_torch(eng)
    nt = len(w.term_off) - 1
    lo = synth.term_at(w.term_bytes, w.term_off, nt // 10)
    hi = synth.term_at(w.term_bytes, w.term_off, nt - nt // 10)
    out = {}
    for root in (0, 1, -1):
        r = eng.read_range_dev(dsegs, lo, hi, None)
        g = eng.read_gather(r, root).download_read()
        out[("read", root)] = (g.n_terms, g.term_bytes, g.term_off, g.post, g.post_off)
    pref = [b"", lo[:1], lo[:2], hi[:2], b"zzzz", b"Q"]
    got = eng.prefix_search_gather(dsegs, pref, 0)
    out["prefix"] = {k: np.asarray(v).tolist() for k, v in got.items()}
    eng.comm_shutdown()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_gather_matches_single_process(orc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from inverted_index_2_b200 import synth
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=500) for _ in range(world))
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    w = _workload()
    nt = len(w.term_off) - 1
    lo = synth.term_at(w.term_bytes, w.term_off, nt // 10)
    hi = synth.term_at(w.term_bytes, w.term_off, nt - nt // 10)
    exp = orc.read_range(w.segments, lo, hi)
    for rank in range(world):
        for root in (0, 1, -1):
            n, tb, toff, post, poff = res[rank][("read", root)]
            if root in (rank, -1):
                assert n == exp.n_terms
                assert np.array_equal(tb, exp.term_bytes) and np.array_equal(toff, exp.term_off)
                assert np.array_equal(post, exp.post) and np.array_equal(poff, exp.post_off)
            else:
                assert n == 0 and toff.tolist() == [0]
    pref = [b"", lo[:1], lo[:2], hi[:2], b"zzzz", b"Q"]
    expp = orc.prefix_search(w.segments, pref)
    assert res[0]["prefix"] == {k: v for k, v in expp.items()}
    assert res[1]["prefix"] == {}
