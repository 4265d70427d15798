#!/bin/bash
T=r02q
V=scratch/variants
timeout 900 python scratch/sweep2.py --libs default:768,$V/libii2_noruns.so:768,$V/libii2_runsonly.so:768,default:768,$V/libii2_noruns.so:768 --steps 10 > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline --steps 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -5 gpurun_out/${T}_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02q_bench.json"))
print(b["ms_per_step"], b.get("range_read_us"), [(k["name"], round(k["ms"]/k["count"],3)) for k in b["kernels"]])
PY
