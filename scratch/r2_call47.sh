#!/bin/bash
# codec kernels as chained launches: parity, C4 at 16 M values (A/B by II2_PDL in separate processes)
T=r05b
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intcomp or val or codec or merge_pipelined" > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
for P in 1 0; do
II2_PDL=$P timeout 600 python bench_extra.py --which c4 --c4-values 16777216 > gpurun_out/${T}_c4_pdl$P.json 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
python - <<PY
import json
b=json.load(open("gpurun_out/${T}_c4_pdl$P.json"))
print("PDL", $P, [(r["L"], r["gap"], round(r["device_encode_ms"],3), round(r["device_decode_ms"],3)) for r in b["results"] if r["codec"]=="intcomp" and r["gap"]==16])
PY
done
