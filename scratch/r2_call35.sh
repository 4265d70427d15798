#!/bin/bash
T=r03g
timeout 900 python scratch/read_small.py > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/${T}_read_launches.csv \
  python scratch/read_small.py --fracs 0.001 --reps 8 > gpurun_out/${T}_ncu.log 2>&1 || tail -3 gpurun_out/${T}_ncu.log
tail -2 gpurun_out/${T}_ncu.log
