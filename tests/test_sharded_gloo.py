"""Multi-rank path on CPU: world_size 2 over gloo, oracle backend (test infrastructure) behind
the host mirror.  Shards are partitioned by shard-key range; compaction needs no collective;
cross-shard reads are gathered in rank order and must equal the single-process index."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from inverted_index_2_b200 import sharded
from host_mirror import InvertedIndex, shard_key
from sharded_harness import ShardedIndex


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _docs():
    rng = np.random.default_rng(4)
    alphabet = b"abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
    vocab = sorted({bytes(rng.choice(list(alphabet), size=int(rng.integers(2, 9))).tolist())
                    for _ in range(400)} | {b"a", b"term1", b"term2"})
    docs = []
    for val in range(1, 41):
        pick = rng.choice(len(vocab), size=30, replace=False)
        docs.append(([vocab[i] for i in pick], val))
    return vocab, docs


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
    from oracle import orc
    from scenario import OracleBackend
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    vocab, docs = _docs()
    weights = np.zeros(sharded.N_SHARD_KEYS)
    for terms, _ in docs:
        for t in terms:
            weights[int(shard_key(t))] += 1
    bounds = sharded.partition_shard_keys(weights, world)
    idx = ShardedIndex(OracleBackend(orc), bounds, rank)
    for terms, val in docs:
        idx.put(terms, val)
    idx.put_removed([3, 7])
    merged = idx.merge(2, 1000)
    full = idx.read(None, None)
    lo, hi = vocab[50], vocab[300]
    part = idx.read(lo, hi)
    pref = idx.prefix_search([b"a", b"te", b"Zz"])
    batched = ShardedIndex(OracleBackend(orc), bounds, rank)
    batched.put_batch(docs)  # one ingest call per shard instead of one Put per document
    full_b = batched.read(None, None)
    if rank == 0:
        q.put((merged, full.items(), part.items(), pref, bounds.tolist(), full_b.items()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_world2_matches_single_process(orc):
    from scenario import OracleBackend
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    merged, full, part, pref, bounds, full_b = q.get(timeout=240)
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    assert 0 < bounds[1] < sharded.N_SHARD_KEYS and merged > 0

    vocab, docs = _docs()
    single = InvertedIndex(OracleBackend(orc))
    for terms, val in docs:
        single.put(terms, val)
    single.put_removed([3, 7])
    single.merge(2, 1000)
    assert full == list(single.read(None, None))
    assert part == list(single.read(vocab[50], vocab[300]))
    assert pref == single.prefix_search([b"a", b"te", b"Zz"])
    plain = InvertedIndex(OracleBackend(orc))
    for terms, val in docs:
        plain.put(terms, val)
    assert full_b == list(plain.read(None, None))
    # removed values are gone after the merge, and the order is the shard-key order
    assert all(3 not in v and 7 not in v for _, v in full)
    keys = [shard_key(t) for t, _ in full]
    assert keys == sorted(keys)


def test_partition_is_contiguous_and_balanced():
    w = np.zeros(sharded.N_SHARD_KEYS)
    w[400:452] = 10  # the synthetic alphabet only reaches 52 shard keys
    for world in (1, 2, 4, 8):
        b = sharded.partition_shard_keys(w, world)
        assert b[0] == 0 and b[-1] == sharded.N_SHARD_KEYS and (np.diff(b) >= 0).all()
        loads = [w[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(loads) - min(loads) <= 2 * 10
        assert sharded.owner_of(425, b) in range(world)
