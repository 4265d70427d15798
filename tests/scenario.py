"""Runs the command sequences of tests/golden/reference_vectors.json against the host
mirror (inverted_index_2_b200.host) with an injected backend — the Python twin of the
reference's TestingMachine (helper_test.go:13-103)."""
from __future__ import annotations

import json
import os

import numpy as np

from host_mirror import InvertedIndex, Shard

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_vectors.json")


def load_vectors() -> dict:
    with open(GOLDEN) as f:
        return json.load(f)


def _b(x):
    return None if x is None else x.encode()


class OracleBackend:
    """Backend protocol of host.py on the CPU oracle (tests only)."""

    def __init__(self, orc):
        self.orc = orc

    def merge(self, segs, removed):
        return self.orc.merge(segs, removed=np.asarray(removed, dtype=np.uint32), decoded=False)

    def read_range(self, segs, min_term, max_term):
        return self.orc.read_range(segs, min_term, max_term)

    def prefix_search(self, segs, prefixes):
        return self.orc.prefix_search(segs, prefixes)

    def ingest(self, docs, removed):
        return self.orc.ingest(docs, removed=np.asarray(removed, dtype=np.uint32), decoded=False)


def run_steps(target, steps, is_index: bool):
    for cmd, arg in steps:
        if cmd == "put":
            target.put([t.encode() for t in arg[0]], arg[1])
        elif cmd == "ingest":
            for val in sorted(arg, key=int):
                target.put([t.encode() for t in arg[val]], int(val))
        elif cmd == "compare":
            expected = sorted((t.encode(), v) for t, v in arg.items())
            assert list(target.read(None, None)) == expected
        elif cmd == "read":
            lo, hi, exp = arg
            assert list(target.read(_b(lo), _b(hi))) == [(t.encode(), v) for t, v in exp]
        elif cmd == "merge":
            req, mx, exp = arg
            got = target.merge(req, mx, 2) if is_index else target.merge(req, mx)
            if exp >= 0:
                assert got == exp
        elif cmd == "remove":
            (target.put_removed if is_index else target.remove)(arg)
        elif cmd == "count_segments":
            assert target.count_segments() == arg
        elif cmd == "removed_values":
            assert target.removed_list.values().tolist() == arg
        elif cmd == "minmax":
            assert target.min_max() == [_b(arg[0]), _b(arg[1])]
        elif cmd == "prefix":
            got = target.prefix_search([p.encode() for p in arg[0]])
            assert got == {k.encode(): v for k, v in arg[1].items()}
        elif cmd == "shards":
            assert len(target.shards) == arg
        else:
            raise ValueError(cmd)


def run_shard_scenario(backend, sc):
    run_steps(Shard("0000", backend), sc["steps"], False)


def run_index_scenario(backend, sc):
    run_steps(InvertedIndex(backend), sc["steps"], True)


# ---- spec-derived roaring byte vectors (tests/golden/roaring_vectors.json) -----------------
def load_roaring_vectors():
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "roaring_vectors.json")) as f:
        return json.load(f)["cases"]


def roaring_case_values(case):
    """The Put() argument of a case, in the order the generator lists it."""
    import numpy as np
    spec = case["values"]
    if "list" in spec:
        return np.array(spec["list"], dtype=np.uint32)
    parts = []
    if "splitmix64" in spec:
        seed, count, mod = spec["splitmix64"]
        state, m64 = seed & 0xFFFFFFFFFFFFFFFF, 0xFFFFFFFFFFFFFFFF
        vals = []
        for _ in range(count):
            state = (state + 0x9E3779B97F4A7C15) & m64
            z = state
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m64
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m64
            vals.append((z ^ (z >> 31)) % mod)
        parts.append(np.array(vals, dtype=np.uint32))
    for key in ("range", "plus_range"):
        if key in spec:
            a, b, step = spec[key]
            parts.append(np.arange(a, b, step, dtype=np.uint32))
    return np.concatenate(parts)


def check_roaring_case(case, data: bytes):
    import hashlib
    assert len(data) == case["len"], (case["name"], len(data), case["len"])
    if "hex" in case:
        assert data.hex() == case["hex"], case["name"]
    else:
        assert data[:64].hex() == case["head_hex"], case["name"]
    assert hashlib.sha256(data).hexdigest() == case["sha256"], case["name"]
