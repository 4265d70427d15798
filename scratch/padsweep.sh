#!/bin/bash
for pad in 0 30000 60000 160000; do
  II2_K1B_PAD=$pad python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/sw.json 2> gpurun_out/sw.err || { echo "RUN FAILED $pad"; tail -2 gpurun_out/sw.err; continue; }
  python - "$pad" <<PY
import json,sys
b=json.load(open("gpurun_out/sw.json"))
print("pad",sys.argv[1], "| ms", round(b["ms_per_step"],3), " ".join("%s=%.3f"%(k["name"][:9],k["ms"]/k["count"]) for k in b["kernels"][1:]))
PY
done
