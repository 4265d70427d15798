#!/bin/bash
# end-of-round evidence (tag = $1): tests, bench lines (default + fused + reference), extra configs,
# ncu launch list and --set full captures of the default pipeline
T=${1:-r02}
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 900 python bench.py --verify > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -5 gpurun_out/${T}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_ref.err || tail -5 gpurun_out/${T}_ref.err
II2_FUSED=1 timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/${T}_bench_fused.json 2> gpurun_out/${T}_bench_fused.err || tail -5 gpurun_out/${T}_bench_fused.err
timeout 900 python bench_extra.py --which c1,c3,prefix > gpurun_out/${T}_bench_extra.json 2> gpurun_out/${T}_extra.err || tail -5 gpurun_out/${T}_extra.err
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-range-read > gpurun_out/${T}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-range-read > gpurun_out/${T}_ncu_l.log 2>&1 || tail -3 gpurun_out/${T}_ncu_l.log
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k1b_group_kernel|k2b_union_kernel|k6_emit_kernel|k1_partition_chunks_raw|k1_bucket_stats|k1_rank_samples' \
  --launch-skip 6 --launch-count 6 -f -o gpurun_out/${T}_full \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-range-read > gpurun_out/${T}_ncu_f.log 2>&1 || tail -3 gpurun_out/${T}_ncu_f.log
python - <<PY
import json
for f in ("${T}_bench","${T}_bench_fused","${T}_bench_reference"):
    try:
        b=json.load(open("gpurun_out/%s.json"%f))
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "value", b.get("value"), "ms", b.get("ms_per_step"), "e2e", (b.get("e2e") or {}).get("ms_per_step"), "range", b.get("range_read_us"), "verified", b.get("verified_full_size"))
    if "kernels" in b: print("   ", [(k["name"], round(k["ms"]/k["count"],3)) for k in b["kernels"]])
PY
ls -la gpurun_out/${T}_*
