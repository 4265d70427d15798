#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intcomp or codec or val" 2>&1 | tail -5
timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c4 or C4 or codec or bitmask" 2>&1 | tail -3
timeout 400 python bench_extra.py --which c4 > gpurun_out/c4.json 2> gpurun_out/c4.err || tail -5 gpurun_out/c4.err
python - <<'PY'
import json
for line in open("gpurun_out/c4.json"):
    d=json.loads(line)
    for r in d["results"]:
        if r["codec"]=="intcomp": print(r["L"],r["lists"],r["gap"],"enc %.0f M/s dec %.0f M/s"%(r["encode_values_per_s"]/1e6,r["decode_values_per_s"]/1e6))
PY
