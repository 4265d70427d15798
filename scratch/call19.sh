#!/bin/bash
timeout 400 python bench_extra.py --which c4 > gpurun_out/c4.json 2> gpurun_out/c4.err || tail -5 gpurun_out/c4.err
python - <<'PY'
import json
for line in open("gpurun_out/c4.json"):
    d=json.loads(line)
    for r in d["results"]:
        if r["codec"]=="intcomp": print(r["L"],r["lists"],r["gap"],"enc %.0f M/s dec %.0f M/s | device enc %.2f ms %.0f GB/s, dec %.2f ms %.0f GB/s"%(r["encode_values_per_s"]/1e6,r["decode_values_per_s"]/1e6,r["device_encode_ms"],r["device_encode_gbs"],r["device_decode_ms"],r["device_decode_gbs"]))
        else: print(r)
PY
