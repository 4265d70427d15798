#!/bin/bash
# small-call bucket size / bucket count after the launch-chain work (one process, env per setting)
T=r06b
timeout 900 python scratch/read_small.py --fracs 0.0002,0.001,0.01 --reps 30 --env "II2_SMALL_BUCKET=128+II2_SMALL_WANT=592;II2_SMALL_BUCKET=64+II2_SMALL_WANT=592;II2_SMALL_BUCKET=256+II2_SMALL_WANT=592;II2_SMALL_BUCKET=512+II2_SMALL_WANT=592;II2_SMALL_BUCKET=128+II2_SMALL_WANT=296;II2_SMALL_BUCKET=128+II2_SMALL_WANT=1184;II2_SMALL_BUCKET=128+II2_SMALL_WANT=592" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
python - <<'PY'
import json
for l in open("gpurun_out/r06b_reads.jsonl"):
    r=json.loads(l); print(r["env"], r["frac"], r["median_us"], r["min_us"])
PY
