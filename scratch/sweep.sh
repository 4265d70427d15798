#!/bin/bash
# usage: sweep.sh "<flags1>" "<flags2>" ...
for f in "$@"; do
  II2_NVCC_EXTRA="$f" python -m inverted_index_2_b200.build > /dev/null 2>&1 || { echo "BUILD FAILED $f"; continue; }
  python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "RUN FAILED $f"; tail -2 gpurun_out/sw.err; continue; }
  python - "$f" <<PY
import json,sys
b=json.load(open("gpurun_out/sw.json"))
print(sys.argv[1], "| ms", round(b["ms_per_step"],3), " ".join("%s=%.3f"%(k["name"][:9],k["ms"]/k["count"]) for k in b["kernels"][1:]))
PY
done
