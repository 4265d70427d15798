"""Term-range sharding of the index over the GPUs of one box (one process per GPU).

The reference already partitions by term prefix into <= 1024 shards that never interact during
Put / Merge (shard.go:19-20, shardKey shard.go:362-378), so compaction needs NO collective:
every rank owns a contiguous range of shard keys and merges its own shards.  Only reads that
span ranks exchange data — InvertedIndex.Read is an ordered concatenation of shard streams
(inverted_index.go:330-338); PrefixSearch adds one sort + compact after the union
(inverted_index.go:274-292).

The exchange itself is the library's: `ii2_comm_init` / `ii2_read_gather` / `ii2_prefix_gather`
(include/ii2.h, csrc/comm.cu — NCCL over NVLink, reachable from cgo).  Protocol: ONE all-gather of
a fixed-size record (sizes / per-prefix offsets), then ONE round of point-to-point transfers of the
flat arrays, unpadded, straight to their place on the root, offsets rebased there.  This module
holds the partitioning rule, thin wrappers over those calls for a `torch.distributed` launch
(`comm_init_from_torch`), and a CPU restatement of the same protocol over a torch process group
(gloo) so that the multi-rank host logic is testable without GPUs.
"""
from __future__ import annotations

import numpy as np

from .flat import ReadResult

N_SHARD_KEYS = 1024  # 10 bits, shard.go:371-375


def shard_key_of(term: bytes) -> int:
    """shardKey (shard.go:362-378) as a number: top 10 bits of the first two bytes; terms
    shorter than two bytes go to shard 0."""
    if len(term) < 2:
        return 0
    return (((term[0] << 8) + term[1]) & 0xFFFF) >> 6


def shard_keys_of_sorted(term_bytes: np.ndarray, term_off: np.ndarray) -> np.ndarray:
    """shardKey of every term of a flat dictionary (vectorised)."""
    off = term_off[:-1].astype(np.int64)
    lens = np.diff(term_off.astype(np.int64))
    tb = np.concatenate([term_bytes, np.zeros(2, dtype=np.uint8)])
    k = ((tb[off].astype(np.int64) << 8) + tb[off + 1].astype(np.int64)) >> 6
    k[lens < 2] = 0
    return k


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


# ---- the library's NCCL exchange under a torch.distributed launch ---------------------------
def comm_init_from_torch(eng, group=None) -> None:
    """ii2_comm_init on every rank of a torch.distributed job: rank 0 makes the NCCL id
    (ii2_comm_unique_id), torch broadcasts its 128 bytes, every rank joins."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [eng.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    eng.comm_init(box[0], rank, world)


# ---- the same protocol on a torch process group (gloo): CPU shim for the tests ------------------
def _exchange(arrays: list[np.ndarray], sizes_all: np.ndarray, root: int, group, device):
    """arrays[j] of every rank -> root, unpadded: per (rank, array) one point-to-point transfer.
    sizes_all[r][j] = elements of array j on rank r (known everywhere from the all-gather).
    Returns on the root a list (per array) of per-rank numpy arrays, elsewhere None."""
    import torch
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    signed = {np.dtype(np.uint8): np.uint8, np.dtype(np.uint32): np.int32, np.dtype(np.uint64): np.int64}
    if rank != root:
        for a in arrays:
            if len(a):  # NCCL / gloo have no unsigned 32/64-bit types: move the bits as signed
                t = torch.from_numpy(np.ascontiguousarray(a).view(signed[a.dtype])).to(device)
                dist.send(t, dst=root, group=group)
        return None
    out = [[None] * world for _ in arrays]
    for r in range(world):
        for j, a in enumerate(arrays):
            n = int(sizes_all[r][j])
            if r == rank:
                out[j][r] = np.ascontiguousarray(a)
            elif n == 0:
                out[j][r] = np.zeros(0, dtype=a.dtype)
            else:
                t = torch.empty(n, dtype=torch.from_numpy(np.zeros(1, dtype=signed[a.dtype])).dtype,
                                device=device)
                dist.recv(t, src=r, group=group)
                out[j][r] = t.cpu().numpy().view(a.dtype)
    return out


def gather_read_results(local: ReadResult, device="cpu", group=None, root: int = 0) -> ReadResult:
    """ii2_read_gather restated: ordered concatenation of every rank's ReadResult on `root`
    (rank order = term order); the other ranks get an empty result."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    T = local.n_terms
    mine = torch.tensor([T, len(local.term_bytes), len(local.post), 0], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(sizes, mine, group=group)  # ONE size record per rank
    sz = np.stack([s.cpu().numpy() for s in sizes])  # [world][4]
    arrays = [np.asarray(local.term_bytes, dtype=np.uint8), np.asarray(local.term_off[:T], dtype=np.uint32),
              np.asarray(local.post, dtype=np.uint32), np.asarray(local.post_off[:T], dtype=np.uint64)]
    per = np.stack([sz[:, 1], sz[:, 0], sz[:, 2], sz[:, 0]], axis=1)
    got = _exchange(arrays, per, root, group, device)
    if got is None:
        return ReadResult(0, np.zeros(0, np.uint8), np.zeros(1, np.uint32), np.zeros(0, np.uint32),
                          np.zeros(1, np.uint64))
    tb_base = np.concatenate([[0], np.cumsum(sz[:, 1])]).astype(np.uint64)
    p_base = np.concatenate([[0], np.cumsum(sz[:, 2])]).astype(np.uint64)
    toff = np.concatenate([(got[1][r].astype(np.uint64) + tb_base[r]).astype(np.uint32) for r in range(world)]
                          + [np.array([tb_base[world]], dtype=np.uint32)])
    poff = np.concatenate([got[3][r] + p_base[r] for r in range(world)]
                          + [np.array([p_base[world]], dtype=np.uint64)])
    return ReadResult(int(sz[:, 0].sum()), np.concatenate(got[0]), toff, np.concatenate(got[2]), poff)


def gather_prefix_results(local: dict, prefixes: list[bytes], device="cpu", group=None, root: int = 0
                          ) -> dict[bytes, list[int]]:
    """ii2_prefix_gather restated: per prefix the sorted-unique union of every rank's values on
    `root` (inverted_index.go:274-292); a prefix is a key iff some rank matched it."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    np_ = len(prefixes)
    voff = np.zeros(np_ + 1, dtype=np.int64)
    matched = np.zeros(np_, dtype=np.int64)
    vals = []
    for i, p in enumerate(prefixes):
        v = np.asarray(local.get(p, []), dtype=np.uint32)
        matched[i] = 1 if p in local else 0
        vals.append(v)
        voff[i + 1] = voff[i] + len(v)
    mine = torch.from_numpy(np.concatenate([voff, matched])).to(device)
    rows = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(rows, mine, group=group)  # fixed size: offsets + matched flags of every rank
    rows = np.stack([r.cpu().numpy() for r in rows])
    flat = np.concatenate(vals) if vals else np.zeros(0, dtype=np.uint32)
    got = _exchange([flat.astype(np.uint32)], rows[:, np_:np_ + 1], root, group, device)
    if got is None:
        return {}
    out: dict[bytes, list[int]] = {}
    for i, p in enumerate(prefixes):
        if not rows[:, np_ + 1 + i].any():
            continue
        parts = [got[0][r][rows[r, i]:rows[r, i + 1]] for r in range(world)]
        out[p] = np.unique(np.concatenate(parts)).astype(np.int64).tolist()  # slices.Sort + Compact
    return out
