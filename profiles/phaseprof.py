import collections, csv, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass,cuda", "--csv",
                      "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
agg = collections.defaultdict(lambda: collections.Counter())
cur, hdr, line = None, None, ''
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur, hdr = r[1].split("/")[-1], None; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if not hdr or len(r) != len(hdr): continue
    if r[0]: line = r[0]
    for name, val in zip(hdr[4:], r[4:]):
        try: agg[(cur, int(line))][name] += float(val)
        except ValueError: pass
phases = eval(sys.argv[3])
out = collections.defaultdict(lambda: collections.Counter())
for (f, l), a in agg.items():
    name = None
    for pf, lo, hi, nm in phases:
        if f == pf and lo <= l < hi: name = nm; break
    if name is None: name = f
    out[name]["smp"] += a["# Samples"]; out[name]["inst"] += a["Instructions Executed"]
    for k, v in a.items():
        if k.startswith("stall_") and "Not Issued" not in k: out[name][k] += v
ts = sum(o["smp"] for o in out.values()); ti = sum(o["inst"] for o in out.values())
for nm, o in sorted(out.items(), key=lambda kv: -kv[1]["smp"]):
    st = sorted(((v, k[6:]) for k, v in o.items() if k.startswith("stall_")), reverse=True)[:3]
    print(f"{o['smp']/ts*100:5.1f}% smp {o['inst']/ti*100:5.1f}% inst  {nm:28s}", " ".join(f"{k}:{v/max(1,o['smp'])*100:.0f}%" for v, k in st))
