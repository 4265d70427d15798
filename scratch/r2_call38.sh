#!/bin/bash
T=r03k
timeout 900 python scratch/read_small.py --fracs 0.000001,0.001 --prof > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
