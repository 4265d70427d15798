#!/bin/bash
# chained launches (programmatic dependent launch) on the merge / read chain: parity, small-read latency, C2
T=r03i
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 900 python scratch/read_small.py --fracs 0.001,0.01,0.1 --env "II2_PDL=1;II2_PDL=0;II2_PDL=1;II2_PDL=0" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
timeout 600 python scratch/sweep2.py --env "II2_PDL=1;II2_PDL=0" --steps 6 > gpurun_out/${T}_c2.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_c2.jsonl
