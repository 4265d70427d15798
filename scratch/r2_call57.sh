#!/bin/bash
# small prefix unions: parity, prefix-search latency
T=r06c
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multirank.py -x -q -m gpu -k "prefix or golden or vectors or scenario or mirror" > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench_extra.py --which prefix > gpurun_out/${T}_prefix.json 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
python - <<'PY'
import json
for l in open("gpurun_out/r06c_prefix.json"):
    r=json.loads(l)
    for x in r["results"]: print(x["prefix_len"], x["prefixes"], x["values_out"], round(x["median_us"],1), x["phase_us"])
PY
