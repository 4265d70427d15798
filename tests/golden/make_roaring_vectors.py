"""Byte vectors for Bitmask.Put (file/bitmask.go:53-59) derived from the PUBLIC RoaringFormatSpec
(https://github.com/RoaringBitmap/RoaringFormatSpec), independently of oracle/roaring_ref.c and of
the CUDA encoder: this script packs the bytes with `struct` straight from the spec's text.

    cookie        no run container: u32 12346, then u32 container count
                  any run container: u32 (12347 | (count-1) << 16), then ceil(count/8) bytes of
                  is-run flags (LSB first)
    descriptive   per container: u16 key, u16 cardinality-1
    offset header u32 byte offset of every container's payload from the start of the stream:
                  always with cookie 12346; with cookie 12347 only when count >= 4
    payloads      array (cardinality <= 4096, not a run): sorted u16 values
                  bitmap: 1024 u64 words
                  run: u16 number of runs, then (u16 start, u16 length-1) per run

Container types are the ones roaring v1.9.4's Add() leaves behind (go.mod:6; restated, ●●○ in
SURVEY.md appendix A.3): array up to 4096 values, bitmap from the 4097th, and a bitmap container
that becomes full (65536) turns into the single run [0, 65535].  RunOptimize is never called.

The dictionary of every case is 0 .. n-1, so index == value (file/bitmask.go:64-71).
Small cases are kept as hex; large ones as length + sha256 + the first 64 bytes.

    python tests/golden/make_roaring_vectors.py   -> tests/golden/roaring_vectors.json
"""
import hashlib
import json
import os
import struct


def spec_serialize(indexes):
    """Portable serialisation of the set `indexes` with the container types Add() produces."""
    by_key = {}
    for x in sorted(set(indexes)):
        by_key.setdefault(x >> 16, []).append(x & 0xFFFF)
    keys = sorted(by_key)
    n = len(keys)
    kinds = []
    for k in keys:
        card = len(by_key[k])
        kinds.append("run" if card == 65536 else ("bitmap" if card > 4096 else "array"))
    has_run = "run" in kinds
    out = bytearray()
    if has_run:
        out += struct.pack("<I", 12347 | ((n - 1) << 16))
        flags = bytearray((n + 7) // 8)
        for i, kind in enumerate(kinds):
            if kind == "run":
                flags[i // 8] |= 1 << (i % 8)
        out += flags
    else:
        out += struct.pack("<II", 12346, n)
    for k in keys:
        out += struct.pack("<HH", k, len(by_key[k]) - 1)
    payloads = []
    for k, kind in zip(keys, kinds):
        vals = by_key[k]
        if kind == "array":
            payloads.append(struct.pack("<%dH" % len(vals), *vals))
        elif kind == "bitmap":
            words = [0] * 1024
            for v in vals:
                words[v >> 6] |= 1 << (v & 63)
            payloads.append(struct.pack("<1024Q", *words))
        else:
            payloads.append(struct.pack("<HHH", 1, 0, 65535))
    if not has_run or n >= 4:
        at = len(out) + 4 * n
        for p in payloads:
            out += struct.pack("<I", at)
            at += len(p)
    for p in payloads:
        out += p
    return bytes(out)


def rng(seed):
    state = seed & 0xFFFFFFFFFFFFFFFF

    def nxt():
        nonlocal state  # splitmix64
        state = (state + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = state
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)
    return nxt


def cases():
    out = []

    def add(name, dict_n, values_spec, values, note):
        b = spec_serialize(values)
        c = {"name": name, "dict_n": dict_n, "values": values_spec, "note": note, "len": len(b),
             "sha256": hashlib.sha256(b).hexdigest()}
        if len(b) <= 256:
            c["hex"] = b.hex()
        else:
            c["head_hex"] = b[:64].hex()
        out.append(c)

    # hand-checkable: 3A30 0000 | 0100 0000 | 0000 0200 | 1000 0000 | 0100 0200 0300
    add("array_one_container", 16, {"list": [3, 1, 2]}, [3, 1, 2],
        "cookie 12346, 1 container, key 0 card-1 2, payload at byte 16, values 1 2 3")
    assert out[-1]["hex"] == "3a300000" "01000000" "00000200" "10000000" "010002000300"
    add("array_with_duplicates_unsorted", 100, {"list": [70, 5, 70, 99, 5, 0]}, [70, 5, 99, 0],
        "Add() is idempotent; array payload is sorted")
    add("array_two_containers", 70000, {"list": [65536, 1, 65537, 69999]}, [65536, 1, 65537, 69999],
        "keys 0 and 1; offsets 24 and 26")
    assert out[-1]["hex"] == ("3a300000" "02000000" "00000000" "01000200" "18000000" "1a000000"
                              "0100" "000001006f11")
    add("array_4096_boundary", 5000, {"range": [0, 4096, 1]}, list(range(4096)),
        "4096 values stay an array container (8192 payload bytes)")
    add("bitmap_4097", 5000, {"range": [0, 4097, 1]}, list(range(4097)),
        "the 4097th value converts the container to a bitmap (8192 payload bytes)")
    add("bitmap_sparse_even", 65536, {"range": [0, 65536, 2]}, list(range(0, 65536, 2)),
        "32768 values: bitmap container of 0x5555... words")
    add("bitmap_65535", 65536, {"range": [0, 65535, 1]}, list(range(65535)),
        "one value short of full: still a bitmap, descriptive header card-1 = 65534")
    add("run_full_container", 65536, {"range": [0, 65536, 1]}, list(range(65536)),
        "full container -> run [0,65535]; cookie 12347, 1 flag byte, NO offset header (count < 4)")
    assert out[-1]["hex"] == "3b300000" "01" "0000ffff" "01000000ffff"
    add("run_then_array_3_containers", 200000, {"range": [0, 131072 + 10, 1]}, list(range(131072 + 10)),
        "2 run containers + 1 array; count 3 < 4: no offset header")
    assert out[-1]["hex"] == ("3b300200" "03" "0000ffff" "0100ffff" "02000900" "01000000ffff"
                              "01000000ffff" + "".join("%02x%02x" % (i, 0) for i in range(10)))
    add("run_with_offset_header_4_containers", 300000, {"range": [0, 3 * 65536 + 5, 1]},
        list(range(3 * 65536 + 5)),
        "3 run containers + 1 array; count 4: offset header present after the descriptive header")
    g = rng(0xB17)
    mixed = sorted({g() % 400000 for _ in range(60000)} | set(range(65536, 131072)))
    add("mixed_array_bitmap_run_7_containers", 400000, {"splitmix64": [0xB17, 60000, 400000],
                                                      "plus_range": [65536, 131072, 1]}, mixed,
        "random 60 000 of 400 000 (bitmap containers) + container 1 filled (run) ")
    g = rng(0xC4)
    sparse = sorted({g() % (1 << 24) for _ in range(3000)})
    add("sparse_256_array_containers", 1 << 24, {"splitmix64": [0xC4, 3000, 1 << 24]}, sparse,
        "3000 values over 2^24: ~256 small array containers, long offset header")
    return out


if __name__ == "__main__":
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "roaring_vectors.json")
    with open(path, "w") as f:
        json.dump({"source": "RoaringFormatSpec (public); container rules of roaring v1.9.4 Add()",
                   "generator": "tests/golden/make_roaring_vectors.py", "cases": cases()}, f, indent=1)
    print(path)
