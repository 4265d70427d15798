#!/usr/bin/env python3
"""Writes tests/golden/reference_vectors.json.

The reference is Go and cannot run in this image (no Go toolchain, dependencies not
vendored), so these vectors are TRANSCRIBED BY HAND from the known answers asserted
in the reference's own tests; each entry cites the file:line it comes from (paths
relative to the reference repo).  They are logical (term -> values) vectors: the
reference holds no byte-level vectors for `_val`, `_fst` or roaring output.

Step vocabulary (mirrors helper_test.go:19-24 TestingMachine commands):
  ["ingest", {val: [terms]}]        IngestBulkCmd  -> Shard.Put(terms, val) per entry
  ["compare", {term: [values]}]     CompareCmd     -> Read(nil,nil) == sorted expected
  ["merge", [req, max, expected]]   MergeCmd       -> Merge(req,max) (expected < 0: any)
  ["remove", [values]]              RemoveCmd      -> Shard.Remove / PutRemoved
  ["count_segments", n]             CountSegmentsCmd
  ["read", [min, max, [[term, [values]], ...]]]    Read(min,max) == expected list
  ["removed_values", [values]]      CheckCmd on removedList.Values()
  ["minmax", [min, max]]            Shard.MinMax()
  ["put", [[terms], val]]           single Put
  ["prefix", [[prefixes], {prefix: [values]}]]     PrefixSearch
  ["shards", n]                     len(ii.shards)
"""
import json
import os

V = {
    "shard_scenarios": [
        {"name": "TestMinMaxTerms", "source": "shard_test.go:16-38", "steps": [
            ["put", [["term1"], 1]], ["minmax", ["term1", "term1"]],
            ["put", [["term2"], 2]], ["minmax", ["term1", "term2"]],
            ["put", [["term1", "term2", "term3"], 3]], ["minmax", ["term1", "term3"]],
        ]},
        {"name": "TestInitFromExistingFiles(content)", "source": "shard_test.go:40-63", "steps": [
            ["put", [["term1", "term2"], 1]], ["put", [["term2", "term3"], 2]],
            ["read", [None, None, [["term1", [1]], ["term2", [1, 2]], ["term3", [2]]]]],
        ]},
        {"name": "TestIngestion", "source": "shard_test.go:65-88", "steps": [
            ["ingest", {"1": ["term1"]}],
            ["compare", {"term1": [1]}],
            ["ingest", {"1": ["term1"], "2": ["term1", "term2"], "3": ["term3"]}],
            ["compare", {"term1": [1, 2], "term2": [2], "term3": [3]}],
        ]},
        {"name": "TestReadPartial(merged)", "source": "shard_test.go:90-136", "steps": [
            ["put", [["AA"], 1]], ["put", [["BB"], 2]], ["put", [["CC"], 3]],
            ["merge", [2, 200, -1]],
            ["read", ["AA", "BB", [["AA", [1]], ["BB", [2]]]]],
            ["read", ["BB", "CC", [["BB", [2]], ["CC", [3]]]]],
        ]},
        {"name": "TestReadPartial(direct)", "source": "shard_test.go:90-136", "steps": [
            ["put", [["AA"], 1]], ["put", [["BB"], 2]], ["put", [["CC"], 3]],
            ["read", ["AA", "BB", [["AA", [1]], ["BB", [2]]]]],
            ["read", ["BB", "CC", [["BB", [2]], ["CC", [3]]]]],
        ]},
        {"name": "TestMerging", "source": "shard_test.go:138-162", "steps": [
            ["ingest", {"1": ["term1"], "2": ["term1", "term2"], "3": ["term3"]}],
            ["count_segments", 3],
            ["merge", [3, 2, 2]], ["count_segments", 2],
            ["merge", [2, 2, 2]], ["count_segments", 1],
            ["merge", [2, 2, 0]], ["count_segments", 1],
            ["compare", {"term1": [1, 2], "term2": [2], "term3": [3]}],
        ]},
        {"name": "TestMergeWithRemoval", "source": "shard_test.go:164-190", "steps": [
            ["ingest", {"1": ["term1", "term3"], "2": ["term2"], "3": ["term3"]}],
            ["count_segments", 3],
            ["merge", [2, 2, 2]], ["count_segments", 2],
            ["remove", [2]],
            ["merge", [2, 2, 2]], ["count_segments", 1],
            ["compare", {"term1": [1], "term3": [1, 3]}],
            ["remove", [10]],
            ["removed_values", [10]],
        ]},
        {"name": "TestMergeEmptySegment", "source": "shard_test.go:192-214", "steps": [
            ["ingest", {"1": ["term1"]}], ["ingest", {"1": ["term1"]}],
            ["remove", [1]],
            ["merge", [2, 2, 2]],
            ["count_segments", 0],
            ["compare", {}],
            ["remove", [2]],
        ]},
        {"name": "TestConcurrentAccess(sequence)", "source": "shard_test.go:216-248", "steps": [
            ["ingest", {"1": ["term1"], "2": ["term1", "term2"], "3": ["term3"]}],
            ["merge", [2, 2, 2]],
            ["compare", {"term1": [1, 2], "term2": [2], "term3": [3]}],
        ]},
    ],
    "index_scenarios": [
        {"name": "TestPutRemove", "source": "inverted_index_test.go:59-82", "steps": [
            ["put", [["aaaa", "bbbb"], 1]], ["put", [["aaaa", "bbbb"], 1]], ["put", [["aaaa"], 2]],
            ["remove", [1]],
            ["merge", [2, 3, -1]],
            ["read", [None, None, [["aaaa", [2]]]]],
        ]},
        {"name": "TestPut", "source": "inverted_index_test.go:140-194", "steps": [
            ["put", [["ab1", "ab2"], 1]], ["put", [["ab2", "cd1"], 2]],
            ["read", [None, None, [["ab1", [1]], ["ab2", [1, 2]], ["cd1", [2]]]]],
            ["shards", 2],
        ]},
        {"name": "TestSearchByPrefix", "source": "inverted_index_test.go:196-220", "steps": [
            ["put", [["a12"], 1]], ["put", [["a13"], 1]], ["put", [["a13"], 2]],
            ["put", [["a20"], 3]], ["put", [["a30"], 4]],
            ["put", [["termA"], 5]], ["put", [["termB"], 6]], ["put", [["termC"], 7]],
            ["prefix", [["a1"], {"a1": [1, 2]}]],
            ["prefix", [["term", "unknown"], {"term": [5, 6, 7]}]],
        ]},
        {"name": "TestReadScoped", "source": "inverted_index_test.go:222-281", "steps": [
            ["put", [["aa"], 1]], ["put", [["bb"], 2]], ["put", [["cc"], 3]], ["put", [["dd"], 4]],
            ["read", [None, None, [["aa", [1]], ["bb", [2]], ["cc", [3]], ["dd", [4]]]]],
            ["read", ["a~", None, [["bb", [2]], ["cc", [3]], ["dd", [4]]]]],
            ["read", [None, "cc", [["aa", [1]], ["bb", [2]], ["cc", [3]]]]],
            ["read", ["bb", "cc", [["bb", [2]], ["cc", [3]]]]],
        ]},
    ],
    # Writer -> Reader round trips, file/writer_test.go
    "writer": [
        {"name": "TestWriter", "source": "file/writer_test.go:11-46", "mode": "val",
         "items": [["term1", [10, 500, 300]], ["term2", []], ["term3", [66, 5513]]]},
        {"name": "TestWriterDirect", "source": "file/writer_test.go:48-84", "mode": "direct",
         "items": [["term1", [10]], ["term2", [11]]]},
    ],
    # file/bitmask_test.go:34-52
    "bitmask": [
        {"name": "TestBitmaskPut", "source": "file/bitmask_test.go:34-52", "init": [],
         "puts": [[1, 10, 80], [9, 10, 11]],
         "get_concat_first": [1, 10, 80],     # Get(v1 || v2) reads only the first bitmap
         "get_second_sorted": [9, 10, 11],    # slices.Sort(Get(v2))
         "get_second_index_order": [10, 9, 11],  # dictionary [1,10,80,9,11] -> indexes 1,3,4
         "all_values": [1, 10, 80, 9, 11]},
    ],
    # removed_list_test.go:9-24
    "removed_lists": {"source": "removed_list_test.go:9-24",
                      "batches": [[1, 5, 10], [2, 20, 30]],
                      "values": [1, 2, 5, 10, 20, 30],
                      "after_sync_second_only": [2, 20, 30]},
    # shardKey, shard.go:362-378 (arithmetic of the function itself) and
    # inverted_index_test.go:140-176 (ab*/cd1 land in 2 shards)
    "shard_key": [["ab1", "0389"], ["ab2", "0389"], ["cd1", "0397"], ["a", "0000"], ["", "0000"],
                  ["aaaa", "0389"], ["bbbb", "0393"]],
}

if __name__ == "__main__":
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")
    with open(out, "w") as f:
        json.dump(V, f, indent=1, sort_keys=True)
    print("wrote", out)
