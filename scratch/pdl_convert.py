#!/usr/bin/env python3
"""One-off source rewrite: make the named kernels of a file chained launches
(pdl_enter() as first statement; `k<<<g, b, m, s>>>(args); II2_LAUNCHED();` -> II2_LAUNCH_CHAIN)."""
import re, sys
path, names = sys.argv[1], sys.argv[2:]
s = open(path).read()

def match_paren(s, i, open_c, close_c):
    d = 0
    while True:
        c = s[i]
        if c == open_c: d += 1
        elif c == close_c:
            d -= 1
            if d == 0: return i
        i += 1

def split_top(a):
    out, d, cur = [], 0, ""
    for c in a:
        if c in "([{<" and not (c == "<" and False): d += c in "([{"
        if c in ")]}": d -= 1
        if c == "," and d == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += c
    out.append(cur.strip())
    return out

for name in names:
    # definition(s)
    nd = 0
    for m in list(re.finditer(r"__global__[^;{]*?\b%s\s*\(" % re.escape(name), s))[::-1]:
        p = match_paren(s, m.end() - 1, "(", ")")
        b = s.index("{", p)
        if s[p + 1:b].strip() != "":  # declaration or something else
            continue
        if s[b + 1:b + 40].lstrip().startswith("pdl_enter();"):
            continue
        s = s[:b + 1] + "\n  pdl_enter();" + s[b + 1:]
        nd += 1
    # launches
    nl = 0
    pat = re.compile(r"\b(%s(?:<[^<>;(){}]*>)?)\s*<<<" % re.escape(name))
    pos = 0
    while True:
        m = pat.search(s, pos)
        if not m: break
        kexpr = m.group(1)
        cfg_start = m.end()
        cfg_end = s.index(">>>", cfg_start)
        cfg = split_top(s[cfg_start:cfg_end])
        while len(cfg) < 4: cfg.append("0")
        a0 = cfg_end + 3
        assert s[a0] == "(", (name, s[a0:a0+20])
        a1 = match_paren(s, a0, "(", ")")
        args = s[a0 + 1:a1]
        rest = s[a1 + 1:]
        mm = re.match(r"\s*;\s*II2_LAUNCHED\(\);", rest)
        assert mm, (name, rest[:60])
        if "," in kexpr: kexpr = "(" + kexpr + ")"
        new = "II2_LAUNCH_CHAIN(%s, %s, %s, %s, %s, %s);" % (kexpr, cfg[0], cfg[1], cfg[2], cfg[3], " ".join(args.split()))
        s = s[:m.start()] + new + s[a1 + 1 + mm.end():]
        pos = m.start() + len(new)
        nl += 1
    print(name, "definitions", nd, "launches", nl)
open(path, "w").write(s)
