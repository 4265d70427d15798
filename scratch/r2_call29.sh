#!/bin/bash
T=r03c
timeout 600 python scratch/sweep2.py --terms 100000 --segments 64 --postings 250000000 --steps 1 --max-len 512 > gpurun_out/${T}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k2_medium_kernel' --launch-skip 2 --launch-count 1 -f -o gpurun_out/${T}_medium \
  python scratch/sweep2.py --terms 100000 --segments 64 --postings 250000000 --steps 1 --max-len 512 > gpurun_out/${T}_ncu.log 2>&1 || tail -5 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_*
