"""CPU restatement of blevesearch/vellum v1.0.10's FST encoding (go.mod:7), independent of
csrc/fst_v1.cpp: a decoder that walks a `<key>_fst` file (decoder_v1.go) and a deliberately
naive encoder that writes the UN-minimised trie (every state its own, no registry) with the
same state encodings (encoder_v1.go) — any vellum reader accepts such a file.

TEST INFRASTRUCTURE ONLY.  PARITY UNPINNED: vellum's source is not in /root/reference and the
reference's tests hold no `_fst` bytes; call sites that fix the semantics: file/writer.go:35,43
(Insert(term, offset | value)), file/reader.go:147-151 (Iterator(min, nil), Current()).
Pure-Python loops: small cases only.
"""
from __future__ import annotations

COMMON_INV = b"te/oasripcnw.hlm-du012g=:bf3y5&_4v9678k%?xCDASFIBEjPTzRNM+LOqHG"
assert len(COMMON_INV) == 63
COMMON = {b: i + 1 for i, b in enumerate(COMMON_INV)}


def _unpack(b: bytes) -> int:
    return int.from_bytes(b, "little")


def _psize(n: int) -> int:
    s = 1
    while n >> (8 * s):
        s += 1
    return s


class State:
    __slots__ = ("final", "final_out", "trans")

    def __init__(self):
        self.final, self.final_out, self.trans = False, 0, []  # trans: (in, out, addr) ascending


def state_at(d: bytes, addr: int) -> State:
    s = State()
    if addr == 0:
        s.final = True
        return s
    hdr = d[addr]
    bottom = addr
    if hdr & 0x80:
        code = hdr & 63
        if code == 0:
            bottom -= 1
            inp = d[bottom]
        else:
            inp = COMMON_INV[code - 1]
        if hdr & 0x40:
            s.trans = [(inp, 0, bottom - 1)]
        else:
            bottom -= 1
            ts, os_ = d[bottom] >> 4, d[bottom] & 15
            bottom -= ts
            delta = _unpack(d[bottom:bottom + ts])
            out = 0
            if os_:
                bottom -= os_
                out = _unpack(d[bottom:bottom + os_])
            s.trans = [(inp, out, bottom - delta if delta else 0)]
        return s
    s.final = bool(hdr & 0x40)
    n = hdr & 63
    if n == 0:
        bottom -= 1
        n = d[bottom]
        if n == 1:
            n = 256
    bottom -= 1
    ts, os_ = d[bottom] >> 4, d[bottom] & 15
    trans_top = bottom
    bottom -= n
    dest_top = bottom
    bottom -= n * ts
    out_top = bottom
    if os_:
        bottom -= n * os_
        if s.final:
            bottom -= os_
            s.final_out = _unpack(d[bottom:bottom + os_])
    for i in range(n):
        inp = d[trans_top - i - 1]
        delta = _unpack(d[dest_top - (i + 1) * ts:dest_top - i * ts])
        out = _unpack(d[out_top - (i + 1) * os_:out_top - i * os_]) if os_ else 0
        s.trans.append((inp, out, bottom - delta if delta else 0))
    return s


def decode(d: bytes) -> list[tuple[bytes, int]]:
    """Every (key, value) in iteration order."""
    assert _unpack(d[0:8]) == 1 and _unpack(d[8:16]) == 0
    n_keys, root = _unpack(d[-16:-8]), _unpack(d[-8:])
    out: list[tuple[bytes, int]] = []

    def walk(addr, key, total):
        s = state_at(d, addr)
        if s.final:
            out.append((bytes(key), total + s.final_out))
        for inp, o, nxt in s.trans:
            key.append(inp)
            walk(nxt, key, total + o)
            key.pop()
    walk(root, bytearray(), 0)
    assert len(out) == n_keys
    return out


def encode_trie(items: list[tuple[bytes, int]]) -> bytes:
    """Un-minimised trie with outputs on the last transition of every key (final outputs 0
    except for the empty key): a valid v1 file that shares no suffixes."""
    class N:
        def __init__(self):
            self.final, self.final_out, self.ch = False, 0, {}
    root = N()
    for k, v in items:
        if not k:
            root.final, root.final_out = True, v
            continue
        n = root
        for b in k[:-1]:
            n = n.ch.setdefault(b, [0, N()])[1]
        e = n.ch.setdefault(k[-1], [0, N()])
        # the value sits on the last edge; deeper keys through it carry their own
        e[1].final = True
        e[1].final_out = v
    buf = bytearray((1).to_bytes(8, "little") + (0).to_bytes(8, "little"))

    def emit(n) -> int:
        tr = [(b, e[0], emit(e[1])) for b, e in sorted(n.ch.items())]
        if not tr and n.final and n.final_out == 0:
            return 0
        start = len(buf)
        deltas = [start - a if a else 0 for _, _, a in tr]
        if len(tr) == 1 and not n.final:
            b, o, a = tr[0]
            osz = _psize(o) if o else 0
            if osz:
                buf.extend(o.to_bytes(osz, "little"))
            tsz = _psize(deltas[0])
            buf.extend(deltas[0].to_bytes(tsz, "little"))
            buf.append(tsz << 4 | osz)
            code = COMMON.get(b, 0)
            if code == 0:
                buf.append(b)
            buf.append(0x80 | code)
            return len(buf) - 1
        anyout = n.final_out != 0 or any(o for _, o, _ in tr)
        osz = max([_psize(n.final_out)] + [_psize(o) for _, o, _ in tr]) if anyout else 0
        tsz = max([_psize(x) for x in deltas], default=0)
        if anyout:
            if n.final:
                buf.extend(n.final_out.to_bytes(osz, "little"))
            for _, o, _ in reversed(tr):
                buf.extend(o.to_bytes(osz, "little"))
        for x in reversed(deltas):
            buf.extend(x.to_bytes(tsz, "little"))
        for b, _, _ in reversed(tr):
            buf.append(b)
        buf.append(tsz << 4 | osz)
        num = len(tr) if len(tr) <= 63 else 0
        if num == 0:
            buf.append(1 if len(tr) == 256 else len(tr))
        buf.append(num | (0x40 if n.final else 0))
        return len(buf) - 1
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    root_addr = emit(root)
    buf.extend(len(items).to_bytes(8, "little"))
    buf.extend(root_addr.to_bytes(8, "little"))
    return bytes(buf)
