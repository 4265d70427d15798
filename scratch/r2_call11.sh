#!/bin/bash
T=r02j
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-range-read > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -5 gpurun_out/${T}_bench.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02j_bench.json"))
print(b["ms_per_step"], [ (k["name"], round(k["ms"]/k["count"],3)) for k in b["kernels"]])
PY
# dense index (1000 postings per term): every term takes the medium path
timeout 600 python scratch/sweep2.py --terms 200000 --segments 64 --postings 200000000 --steps 3 > gpurun_out/${T}_dense.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense.jsonl
timeout 900 python bench_extra.py --which c4 > gpurun_out/${T}_c4.json 2> gpurun_out/${T}_c4.err || tail -5 gpurun_out/${T}_c4.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02j_c4.json"))
for r in b["results"]:
    if r["codec"]=="intcomp":
        print("intcomp L",r["L"],"gap",r["gap"],"enc GB/s",round(r["device_encode_gbs"] or 0),"dec GB/s",round(r["device_decode_gbs"] or 0), "frac", round(r.get("encode_frac_of_hbm_peak") or 0,3), round(r.get("decode_frac_of_hbm_peak") or 0,3))
    else:
        print("bitmask L",r["L"],"put ms",r["device_put_ms"],"put Gv/s", round((r.get("device_put_values_per_s") or 0)/1e9,2),"get ms",r["device_get_ms"])
PY
