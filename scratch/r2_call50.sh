#!/bin/bash
# last sanity of the round: whole gpu suite, smoke, the default bench line
T=r05
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -5 gpurun_out/${T}_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r05_bench.json"))
print("value", b["value"], "ms", b["ms_per_step"], "e2e", b["e2e"]["ms_per_step"], "range", b.get("range_read_us"), "launches", b["gpu_launches"], "roof", b["roofline"]["frac"])
PY
