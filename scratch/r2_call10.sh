#!/bin/bash
T=r02i
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-range-read > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -5 gpurun_out/${T}_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02i_bench.json"))
print(b["ms_per_step"], [ (k["name"], round(k["ms"]/k["count"],3)) for k in b["kernels"]])
PY
bash scratch/r2_call9.sh memcheck
