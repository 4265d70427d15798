#!/usr/bin/env python3
"""The workload of the compute-sanitizer passes (profiles/r02_sanitizer_*.log): smoke() plus the
parity tests with the trickiest kernels, on both bucket pipelines.  Small inputs: the tools slow
every kernel 10-50x.   usage: compute-sanitizer --tool memcheck python scratch/sanitize_run.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import __graft_entry__ as entry  # noqa: E402
from inverted_index_2_b200.engine import Engine  # noqa: E402
from oracle import orc  # noqa: E402

import test_gpu_parity as T  # noqa: E402

eng = Engine.default()
orc.lib()
for path in ("general", "fused"):
    if path == "fused":
        os.environ["II2_FUSED"] = "1"
    else:
        os.environ.pop("II2_FUSED", None)
    entry.smoke()
    T.test_union_width_boundaries(eng, orc, path)
    T.test_heavy_terms_multi_cta_union(eng, orc, path)
    T.test_merge_edge_cases(eng, orc, path)
    T.test_read_range_matches_oracle(eng, orc, path)
    print("path", path, "ok", flush=True)
os.environ.pop("II2_FUSED", None)
T.test_intcomp_long_lists_block_parallel(eng, orc)
T.test_bitmask_full_chunk_run_container(eng, orc)
T.test_comm_world1_read_and_prefix_gather(eng, orc)
print("sanitize_run ok", flush=True)
