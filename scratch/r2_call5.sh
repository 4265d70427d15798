#!/bin/bash
T=r02e
export II2_COALESCE=1
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k12f_bucket_kernel' --launch-skip 1 --launch-count 1 -f -o gpurun_out/${T}_full \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1 || tail -5 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_*
