#!/bin/bash
T=r02g
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -8 gpurun_out/${T}_tests.log
timeout 900 python bench.py --verify --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -15 gpurun_out/${T}_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02g_bench.json"))
for k in ("value","ms_per_step","e2e","e2e_decoded","range_read_us","roofline","pipeline","verified_full_size","cpu_baseline"):
    print(k, json.dumps(b.get(k))[:400])
print([ (k["name"], round(k["ms"]/k["count"],3)) for k in b["kernels"]])
PY
