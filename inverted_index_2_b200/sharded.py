"""Term-range sharding of the index over the GPUs of one box (one process per GPU).

The reference already partitions by term prefix into <= 1024 shards that never interact during
Put / Merge (shard.go:19-20, shardKey shard.go:362-378), so compaction needs NO collective:
every rank owns a contiguous range of shard keys and merges its own shards.  Only reads that
span ranks exchange data — InvertedIndex.Read is an ordered concatenation of shard streams
(inverted_index.go:330-338), so the gather is: all-gather of the sizes, then a padded
all-gather of the flat arrays, concatenated in rank order (rank order == shard-key order).
PrefixSearch adds one dedup pass after the gather (inverted_index.go:289-292).
`torch.distributed` is the plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np

from .flat import ReadResult
from .host import InvertedIndex, shard_key

N_SHARD_KEYS = 1024  # 10 bits, shard.go:371-375


def partition_shard_keys(weights: np.ndarray, world: int) -> np.ndarray:
    """Contiguous shard-key ranges balanced by `weights` (e.g. postings per shard key).
    Returns bounds[world+1]: rank r owns shard keys [bounds[r], bounds[r+1])."""
    w = np.asarray(weights, dtype=np.float64)
    assert len(w) == N_SHARD_KEYS
    cum = np.concatenate([[0.0], np.cumsum(w)])
    total = cum[-1]
    bounds = np.zeros(world + 1, dtype=np.int64)
    bounds[world] = N_SHARD_KEYS
    for r in range(1, world):
        bounds[r] = int(np.searchsorted(cum, total * r / world, side="left"))
    return np.maximum.accumulate(bounds)


def owner_of(key: int, bounds: np.ndarray) -> int:
    return int(np.searchsorted(bounds, key, side="right") - 1)


def _gather_var(dist, t, group=None):
    """All-gather of 1-D tensors of different lengths; returns the list in rank order."""
    import torch
    world = dist.get_world_size(group)
    n = torch.tensor([t.numel()], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def gather_read_results(local: ReadResult, device="cpu", group=None) -> ReadResult:
    """Ordered concatenation of every rank's ReadResult (rank order = term order)."""
    import torch
    import torch.distributed as dist

    def tt(a, dt):
        # NCCL has no unsigned 32/64-bit types: move the bits as signed integers
        return torch.from_numpy(np.ascontiguousarray(a).view(dt)).to(device)
    tb = _gather_var(dist, tt(local.term_bytes, np.uint8), group)
    toff = _gather_var(dist, tt(local.term_off, np.int32), group)
    post = _gather_var(dist, tt(local.post, np.int32), group)
    poff = _gather_var(dist, tt(local.post_off, np.int64), group)
    term_bytes = torch.cat(tb).cpu().numpy()
    posts = torch.cat(post).cpu().numpy().view(np.uint32)
    t_parts, p_parts = [np.zeros(1, dtype=np.uint32)], [np.zeros(1, dtype=np.uint64)]
    tbase, pbase = 0, 0
    n_terms = 0
    for to, po in zip(toff, poff):
        to = to.cpu().numpy().view(np.uint32).astype(np.uint64)
        po = po.cpu().numpy().view(np.uint64)
        if len(to) > 1:
            t_parts.append((to[1:] + tbase).astype(np.uint32))
            p_parts.append(po[1:] + np.uint64(pbase))
            n_terms += len(to) - 1
        tbase += int(to[-1]) if len(to) else 0
        pbase += int(po[-1]) if len(po) else 0
    return ReadResult(n_terms, term_bytes, np.concatenate(t_parts), posts, np.concatenate(p_parts))


class ShardedIndex:
    """One rank's slice of the index: the shards whose key falls in its range."""

    def __init__(self, backend, bounds: np.ndarray, rank: int, device="cpu", group=None):
        self.local = InvertedIndex(backend)
        self.bounds, self.rank, self.device, self.group = bounds, rank, device, group

    def _mine(self, term: bytes) -> bool:
        return owner_of(int(shard_key(term)), self.bounds) == self.rank

    def put(self, terms: list[bytes], val: int) -> None:
        """Every rank sees the Put; each keeps the terms of its own shards."""
        mine = [t for t in terms if self._mine(t)]
        if mine:
            self.local.put(mine, val)

    def put_batch(self, docs: list[tuple[list[bytes], int]]) -> None:
        """Batched ingest (ii2_ingest per shard): every rank keeps the terms of its own shards."""
        mine = [([t for t in terms if self._mine(t)], val) for terms, val in docs]
        self.local.put_batch([(t, v) for t, v in mine if t])

    def put_removed(self, values) -> None:
        self.local.put_removed(values)  # tombstones go to every shard (inverted_index.go:41-55)

    def merge(self, req_count: int, m_count: int) -> int:
        return self.local.merge(req_count, m_count)  # shards are independent: no collective

    def _local_read(self, min_term, max_term) -> ReadResult:
        items = list(self.local.read(min_term, max_term))
        from .flat import FlatSegment
        seg = FlatSegment.from_items(items)
        return ReadResult(len(items), seg.term_bytes, seg.term_off, seg.post, seg.post_off)

    def read(self, min_term: bytes | None = None, max_term: bytes | None = None) -> ReadResult:
        """Cross-shard Read: local ordered read, then the NCCL/gloo gather."""
        return gather_read_results(self._local_read(min_term, max_term), self.device, self.group)

    def prefix_search(self, prefixes: list[bytes]) -> dict[bytes, list[int]]:
        """Per-rank prefix search, gathered, then the final sort + compact
        (inverted_index.go:289-292)."""
        import torch.distributed as dist
        local = self.local.prefix_search(prefixes)
        world = dist.get_world_size(self.group)
        parts = [None] * world
        dist.all_gather_object(parts, local, group=self.group)
        out: dict[bytes, list[int]] = {}
        for p in parts:
            for k, v in p.items():
                out.setdefault(k, []).extend(v)
        return {k: sorted(set(v)) for k, v in out.items()}
