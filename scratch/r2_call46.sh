#!/bin/bash
# upload stream at the highest priority: e2e A/B (separate processes: the stream is created once)
T=r05a
for P in 1 0 1 0; do
II2_AUX_PRIORITY=$P timeout 600 python bench.py --no-cpu-baseline --no-range-read --steps 5 --warmup 3 > gpurun_out/${T}_e2e_p$P.json 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
python - <<PY
import json
b=json.load(open("gpurun_out/${T}_e2e_p$P.json"))
print("priority", $P, "e2e ms", b["e2e"]["ms_per_step"], "value ms", b["ms_per_step"], "decoded e2e", (b.get("e2e_decoded") or {}).get("ms_per_step"))
PY
done
