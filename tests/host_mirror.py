"""TEST HARNESS — host-side mirror of the reference's public surface for the hot path (it replays
the reference's own test sequences against the C-ABI; SURVEY §2 rows 8, 10, 11 are out of scope
for the product and stay in Go, so this file lives under tests/ and is not shipped).

Same names, argument meaning and error behaviour as the Go API —
InvertedIndex.{Put, PutRemoved, Merge, Read, PrefixSearch} (inverted_index.go:41,
62,113,192,300), Shard.{Put, Read, Remove, Merge, MinMax} (shard.go:33,72,78,127,
280), Segments (segments.go), RemovedLists (removed_list.go) — with the merge loop
(shard.go:158-212) and the read iterator (shard.go:253-278) replaced by ONE call
each into the C-ABI (include/ii2.h).  Segments are held as flat views
(flat.FlatSegment); the vellum `_fst` file itself stays with the Go host (see
INTEGRATION.md), so "a segment" here is what Go hands over after iterating its FST.

The compute backend is injected: the product default is the CUDA engine
(engine.Engine); tests may inject the CPU oracle to check this host logic without a
GPU.  There is no built-in CPU path.
"""
from __future__ import annotations

import bisect
import os
import threading
import time
from typing import Iterator, Protocol

import numpy as np

from inverted_index_2_b200 import _abi as A
from inverted_index_2_b200.flat import FlatSegment, MergeResult, ReadResult


class Backend(Protocol):
    def merge(self, segs: list[FlatSegment], removed: np.ndarray) -> MergeResult: ...

    def read_range(self, segs: list[FlatSegment], min_term: bytes | None,
                   max_term: bytes | None) -> ReadResult: ...

    # optional: found[prefix] (sorted-unique values) for every prefix with a matching term
    # def prefix_search(self, segs: list[FlatSegment], prefixes: list[bytes]) -> dict: ...


_key_lock = threading.Lock()
_last_key = 0


def _unix_nano_key() -> int:
    """Segment key = UnixNano (file/writer.go:98); kept strictly increasing."""
    global _last_key
    with _key_lock:
        k = time.time_ns()
        if k <= _last_key:
            k = _last_key + 1
        _last_key = k
        return k


def shard_key(term: bytes) -> str:
    """shardKey, shard.go:362-378."""
    if len(term) < 2:
        term = b"\x00\x00"
    return "%04d" % ((((term[0] << 8) + term[1]) & 0xFFFF) >> 6)


class RemovedLists:
    """removed_list.go:14-80 (gob persistence is the Go host's, out of scope)."""

    def __init__(self):
        self.lists: dict[int, np.ndarray] = {}
        self.m = threading.RLock()

    def put(self, timestamp: int, values) -> None:
        with self.m:
            self.lists[timestamp] = np.asarray(values, dtype=np.uint32)

    def values(self) -> np.ndarray:
        """All batches concatenated and sorted, duplicates kept (:44-54)."""
        with self.m:
            if not self.lists:
                return np.zeros(0, dtype=np.uint32)
            return np.sort(np.concatenate(list(self.lists.values())))

    def sync(self, timestamps: list[int]) -> None:
        if not timestamps:
            return
        with self.m:
            oldest = min(timestamps)
            for t in [t for t in self.lists if t < oldest]:
                del self.lists[t]

    def serialize(self) -> bytes:
        """RemovedLists.Serialize, removed_list.go:73-80 (gob; csrc/removed_gob.cpp)."""
        from inverted_index_2_b200 import fst
        with self.m:
            return fst.removed_list_encode(self.lists)

    @classmethod
    def unserialize(cls, data: bytes) -> "RemovedLists":
        """UnserializeRemovedList, removed_list.go:26-33."""
        from inverted_index_2_b200 import fst
        rl = cls()
        rl.lists = fst.removed_list_decode(data)
        return rl


class Segment:
    """segments.go:16-23."""

    def __init__(self, key: int, terms: int, min_term: bytes | None, max_term: bytes | None,
                 data: FlatSegment):
        self.key = key
        self.terms = terms
        self.min_term = min_term
        self.max_term = max_term
        self.data = data
        self.merging = False
        self.readers = 0  # stand-in for the per-segment RWMutex reader count


class Segments:
    """segments.go:10-93: list ordered by term count (smallest first, for merging)."""

    def __init__(self):
        self.list: list[Segment] = []
        self.m = threading.RLock()

    def add(self, seg: Segment) -> None:
        with self.m:
            pos = bisect.bisect_left([s.terms for s in self.list], seg.terms)
            self.list.insert(pos, seg)

    def __len__(self) -> int:
        with self.m:
            return len(self.list)

    def read_lock_all(self) -> list[Segment]:
        with self.m:
            for s in self.list:
                s.readers += 1
            return list(self.list)

    def read_release(self, segs: list[Segment]) -> None:
        with self.m:
            for s in segs:
                s.readers -= 1

    def detach(self, segs: list[Segment]) -> None:
        with self.m:
            ids = {id(s) for s in segs}
            self.list = [s for s in self.list if id(s) not in ids]


class Shard:
    """shard.go: one term-prefix shard = a set of immutable segments + removed list."""

    def __init__(self, key: str, backend: Backend, basedir: str | None = None):
        """basedir: the shard's directory of `<key>_fst` / `<key>_val` files and `removed.list`
        (NewShard, shard.go:300-358); existing segments and the removed list are loaded, new
        ones are written there.  None keeps the shard in memory only."""
        self.key = key
        self.backend = backend
        self.basedir = basedir
        self.segments = Segments()
        self.removed_list = RemovedLists()
        if basedir is not None:
            from inverted_index_2_b200 import files
            os.makedirs(basedir, exist_ok=True)
            for k in files.list_segments(basedir):  # shard.go:307-331
                data = files.open_segment(basedir, k)
                if data is None:
                    continue
                terms = data.terms()
                self.segments.add(Segment(int(k), len(terms), terms[0], terms[-1], data))
            rl = os.path.join(basedir, "removed.list")  # shard.go:340-352
            if os.path.exists(rl):
                with open(rl, "rb") as f:
                    self.removed_list = RemovedLists.unserialize(f.read())

    def _persist_removed(self) -> None:
        """The removed list is flushed with every change (Shard.Remove, shard.go:95-104)."""
        if self.basedir is not None:
            tmp = os.path.join(self.basedir, "removed.list.tmp")
            with open(tmp, "wb") as f:
                f.write(self.removed_list.serialize())
            os.rename(tmp, os.path.join(self.basedir, "removed.list"))

    def _persist(self, seg: "Segment") -> None:
        if self.basedir is not None:
            from inverted_index_2_b200 import files
            files.write_segment(self.basedir, str(seg.key), seg.data)

    def _unlink(self, segs: list["Segment"]) -> None:
        if self.basedir is not None:
            from inverted_index_2_b200 import files
            for s in segs:
                files.remove_segment(self.basedir, str(s.key))

    def get_key(self) -> str:
        return self.key

    def put(self, terms: list[bytes], val: int) -> None:
        """Shard.Put, shard.go:33-67: one direct-mode segment per ingested document."""
        terms = sorted(terms)
        data = FlatSegment.direct(terms, val)
        seg = Segment(_unix_nano_key(), len(terms), terms[0] if terms else None,
                      terms[-1] if terms else None, data)
        self._persist(seg)
        self.segments.add(seg)

    def put_batch(self, docs: list[tuple[list[bytes], int]]) -> None:
        """Extension: Put for every document of the batch followed by one
        Merge of exactly those documents, as ONE C-ABI call (ii2_ingest) that leaves one full
        segment instead of len(docs) direct-mode ones.  Needs a backend with `ingest`."""
        docs = [(list(t), int(v)) for t, v in docs if len(t)]
        for i in range(0, len(docs), 1024):
            chunk = docs[i:i + 1024]
            res = self.backend.ingest(chunk, self.removed_list.values())
            if res.terms_count > 0:
                seg = Segment(_unix_nano_key(), res.terms_count, res.min_term, res.max_term,
                              res.to_segment())
                self._persist(seg)
                self.segments.add(seg)

    def read(self, min_term: bytes | None = None, max_term: bytes | None = None
             ) -> Iterator[tuple[bytes, list[int]]]:
        """Shard.Read, shard.go:72-75: union over ALL segments, [min,max] inclusive.
        No removed filter on reads (survey Q2)."""
        segs = self.segments.read_lock_all()
        try:
            res = self.backend.read_range([s.data for s in segs], min_term, max_term)
        finally:
            self.segments.read_release(segs)
        return iter(res.items())

    def remove(self, values) -> None:
        """Shard.Remove, shard.go:78-105."""
        values = list(values)
        if not values:
            return
        now = _unix_nano_key()
        with self.segments.m:
            stamps = [now] + [s.key for s in self.segments.list]
        self.removed_list.sync(stamps)
        self.removed_list.put(_unix_nano_key(), values)
        self._persist_removed()

    def merge(self, req_count: int, m_count: int) -> int:
        """Shard.Merge, shard.go:127-245; returns how many segments were merged."""
        if len(self.segments) < req_count:
            return 0
        chosen: list[Segment] = []
        with self.segments.m:
            for s in self.segments.list:  # ascending term count
                if len(chosen) == m_count:
                    break
                if not s.merging:  # CompareAndSwap(false, true), :141
                    s.merging = True
                    chosen.append(s)
        if len(chosen) < 2:
            return 0  # NB: like the reference, the claimed flag is not cleared (Q8)
        removed = self.removed_list.values()
        res = self.backend.merge([s.data for s in chosen], removed)
        if res.terms_count > 0:  # lazy writer: nothing is written for an empty result
            seg = Segment(_unix_nano_key(), res.terms_count, res.min_term, res.max_term,
                          res.to_segment())
            self._persist(seg)
            self.segments.add(seg)
        self.segments.detach(chosen)
        self._unlink(chosen)  # removeSegments, shard.go:232-242
        return len(chosen)

    def min_max(self) -> list[bytes | None]:
        """Shard.MinMax, shard.go:280-298."""
        lo = hi = None
        with self.segments.m:
            for s in self.segments.list:
                if lo is None or (s.min_term is not None and lo > s.min_term):
                    lo = s.min_term
                if hi is None or (s.max_term is not None and hi < s.max_term):
                    hi = s.max_term
        return [lo, hi]

    def count_segments(self) -> int:
        return len(self.segments)


class InvertedIndex:
    """inverted_index.go: router over term-prefix shards."""

    def __init__(self, backend: Backend | None = None, basedir: str | None = None):
        """basedir: one sub-directory per shard key, each holding that shard's segment files
        (NewInvertedIndex, inverted_index.go:342-378); existing shards are loaded."""
        if backend is None:
            from inverted_index_2_b200.engine import Engine  # the CUDA engine; raises if unavailable
            backend = Engine.default()
        self.backend = backend
        self.basedir = basedir
        self.shards: list[Shard] = []  # sorted by key
        self.m = threading.RLock()
        if basedir is not None:
            os.makedirs(basedir, exist_ok=True)
            for name in sorted(os.listdir(basedir)):
                if os.path.isdir(os.path.join(basedir, name)):
                    self.shards.append(Shard(name, backend, os.path.join(basedir, name)))

    def _find_shard(self, key: str) -> Shard | None:
        with self.m:
            keys = [s.key for s in self.shards]
            i = bisect.bisect_left(keys, key)
            return self.shards[i] if i < len(keys) and keys[i] == key else None

    def _new_shard(self, key: str) -> Shard:
        with self.m:
            keys = [s.key for s in self.shards]
            i = bisect.bisect_left(keys, key)
            if i < len(keys) and keys[i] == key:
                return self.shards[i]
            sh = Shard(key, self.backend,
                       os.path.join(self.basedir, key) if self.basedir is not None else None)
            self.shards.insert(i, sh)
            return sh

    def put(self, terms: list[bytes], val: int) -> None:
        """InvertedIndex.Put, inverted_index.go:113-145."""
        groups: dict[str, list[bytes]] = {}
        for t in terms:
            groups.setdefault(shard_key(t), []).append(t)
        for key in sorted(groups):
            shard = self._find_shard(key) or self._new_shard(key)
            shard.put(groups[key], val)

    def put_batch(self, docs: list[tuple[list[bytes], int]]) -> None:
        """Extension: a batch of documents routed by shard key, one ii2_ingest per shard."""
        per_shard: dict[str, list[tuple[list[bytes], int]]] = {}
        for terms, val in docs:
            groups: dict[str, list[bytes]] = {}
            for t in terms:
                groups.setdefault(shard_key(t), []).append(t)
            for key, ts in groups.items():
                per_shard.setdefault(key, []).append((ts, val))
        for key in sorted(per_shard):
            shard = self._find_shard(key) or self._new_shard(key)
            shard.put_batch(per_shard[key])

    def put_removed(self, values) -> None:
        """InvertedIndex.PutRemoved, inverted_index.go:41-55: every shard gets the batch."""
        with self.m:
            shards = list(self.shards)
        for sh in shards:
            sh.remove(values)

    def merge(self, req_count: int, m_count: int, concurrency: int = 1) -> int:
        """InvertedIndex.Merge, inverted_index.go:62-109.  Shards are independent; on the
        GPU each shard.Merge is one coarse C-ABI call, `concurrency` host threads."""
        with self.m:
            shards = list(self.shards)
        total = 0
        if concurrency <= 1:
            for sh in shards:
                total += sh.merge(req_count, m_count)
            return total
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(concurrency) as ex:
            for n in ex.map(lambda sh: sh.merge(req_count, m_count), shards):
                total += n
        return total

    def read(self, min_term: bytes | None = None, max_term: bytes | None = None
             ) -> Iterator[tuple[bytes, list[int]]]:
        """InvertedIndex.Read, inverted_index.go:300-340: shards whose [min,max] overlap
        the range, concatenated in shard-key order (lazily, one shard at a time)."""
        with self.m:
            shards = list(self.shards)
        picked = []
        for s in shards:
            lo, hi = s.min_max()
            if lo is None:  # shard without segments
                continue
            if min_term is not None and min_term > hi:
                continue
            if max_term is not None and max_term < lo:
                continue
            picked.append(s)

        def gen():
            for s in picked:
                yield from s.read(min_term, max_term)
        return gen()

    def prefix_search(self, prefixes: list[bytes]) -> dict[bytes, list[int]]:
        """InvertedIndex.PrefixSearch, inverted_index.go:192-295.  Shard selection by min/max
        as the reference does it (:211-236); the scan of the selected shards (:239-285) and the
        final sort + compact (:289-292) are ONE C-ABI call (ii2_prefix_search) over the
        segments of all selected shards when the backend offers it."""
        prefixes = sorted(prefixes)
        with self.m:
            shards = list(self.shards)
        picked: list[tuple[Shard, list[bytes]]] = []
        for shard in shards:
            lo, hi = shard.min_max()
            if lo is None:
                continue
            mine = []
            for p in prefixes:
                l = min(len(p), len(lo))
                if p[:l] < lo[:l]:
                    continue
                l = min(len(p), len(hi))
                if p[:l] > hi[:l]:
                    continue
                mine.append(p)
            if mine:
                picked.append((shard, mine))
        if not picked:
            return {}
        if hasattr(self.backend, "prefix_search"):
            wanted = sorted({p for _, mine in picked for p in mine})
            locked = [(shard, shard.segments.read_lock_all()) for shard, _ in picked]
            try:
                segs = [s.data for _, ss in locked for s in ss]
                res = self.backend.prefix_search(segs, wanted)
            finally:
                for shard, ss in locked:
                    shard.segments.read_release(ss)
            return {k: [int(x) for x in v] for k, v in res.items()}
        found: dict[bytes, list[int]] = {}
        for shard, mine in picked:  # backend without the fused call: scan through Shard.Read
            greatest = mine[-1]
            for term, values in shard.read(mine[0], None):
                if greatest < term[:min(len(term), len(greatest))]:
                    break
                for p in mine:
                    if term.startswith(p):
                        found.setdefault(p, []).extend(values)
        return {k: sorted(set(v)) for k, v in found.items()}  # slices.Sort + Compact, :289-292
