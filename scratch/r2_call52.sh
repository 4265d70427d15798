#!/bin/bash
# final captures: per-term kernels (ncu --set full), kernel sequence of a point read and a 0.1 % read
T=r05f
timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k2_mwarp_kernel|k2_medium_kernel' --launch-skip 6 --launch-count 3 -f -o gpurun_out/${T}_perterm \
  python scratch/sweep2.py --terms 200000 --segments 64 --postings 200000000 --steps 1 --max-len 512 > gpurun_out/${T}_ncu.log 2>&1 || tail -5 gpurun_out/${T}_ncu.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${T}_read_launches.csv \
  python scratch/read_small.py --fracs 0.000001,0.001 --reps 7 > gpurun_out/${T}_ncu2.log 2>&1 || tail -3 gpurun_out/${T}_ncu2.log
ls -la gpurun_out/${T}_*
