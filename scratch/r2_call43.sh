#!/bin/bash
# C3 (256 segments): bucket size against the per-bucket header (256 runs per bucket)
T=r04y
timeout 900 python scratch/read_small.py --fracs 0.1,1.0 --reps 12 --env "II2_BUCKET=576;II2_BUCKET=768;II2_BUCKET=1024;II2_BUCKET=576" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
