"""TEST HARNESS — one rank's slice of a sharded index on top of the host mirror
(tests/host_mirror.py): routes Put / PutRemoved / Merge to the shards whose key falls in the rank's
range and gathers cross-shard reads with the exchange protocol of inverted_index_2_b200/sharded.py
(gloo on the CPU; on GPUs the library's own ii2_read_gather / ii2_prefix_gather)."""
from __future__ import annotations

import numpy as np

from host_mirror import InvertedIndex
from inverted_index_2_b200.flat import FlatSegment, ReadResult
from inverted_index_2_b200.sharded import (gather_prefix_results, gather_read_results, owner_of,
                                           shard_key_of)


class ShardedIndex:
    """One rank's slice of the index: the shards whose key falls in its range."""

    def __init__(self, backend, bounds: np.ndarray, rank: int, device="cpu", group=None):
        self.local = InvertedIndex(backend)
        self.bounds, self.rank, self.device, self.group = bounds, rank, device, group

    def _mine(self, term: bytes) -> bool:
        return owner_of(shard_key_of(term), self.bounds) == self.rank

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
        seg = FlatSegment.from_items(items)
        return ReadResult(len(items), seg.term_bytes, seg.term_off, seg.post, seg.post_off)

    def read(self, min_term: bytes | None = None, max_term: bytes | None = None, root: int = 0
             ) -> ReadResult:
        """Cross-shard Read: local ordered read, then the gather to `root`."""
        return gather_read_results(self._local_read(min_term, max_term), self.device, self.group, root)

    def prefix_search(self, prefixes: list[bytes], root: int = 0) -> dict[bytes, list[int]]:
        """Per-rank prefix search, gathered to `root`, then the final sort + compact
        (inverted_index.go:289-292)."""
        return gather_prefix_results(self.local.prefix_search(prefixes), prefixes, self.device,
                                     self.group, root)
