#!/usr/bin/env python3
"""Aggregate an ncu source page (sass,cuda view) by CUDA source line.
usage: srcprof.py report.ncu-rep kernel_regex [top]"""
import collections
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
SORTKEY = "Instructions Executed" if len(sys.argv) > 4 else "# Samples"
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.defaultdict(lambda: collections.Counter())
src = {}
cur, hdr, line, text = None, None, '', ''
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if not hdr or len(r) != len(hdr):
        continue
    if r[0]:
        line, text = r[0], r[1].strip()
    key = (cur, line)
    src[key] = text
    for name, val in zip(hdr[4:], r[4:]):
        try:
            agg[key][name] += float(val)
        except ValueError:
            pass
tot = sum(a["# Samples"] for a in agg.values()) or 1
toti = sum(a["Instructions Executed"] for a in agg.values()) or 1
print(f"samples {tot:.0f}  warp-instructions {toti:.0f}")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][SORTKEY])[:top]:
    stalls = sorted(((v, k) for k, v in a.items() if k.startswith("stall_") and "Not Issued" not in k),
                    reverse=True)[:2]
    st = " ".join(f"{k[6:]}:{v / max(1, a['# Samples']) * 100:.0f}%" for v, k in stalls)
    print(f"{a['# Samples'] / tot * 100:5.1f}% smp {a['Instructions Executed'] / toti * 100:5.1f}% inst "
          f"{key[0]}:{key[1]:>4} {src[key][:64]:64s} {st}")
