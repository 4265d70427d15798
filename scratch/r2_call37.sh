#!/bin/bash
# single-bucket plan for point reads: parity, latency of 1-term / 20-term / 0.1 % reads
T=r03j
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 900 python scratch/read_small.py --fracs 0.000001,0.00002,0.0001,0.001 --env "II2_SINGLE_BUCKET=768;II2_SINGLE_BUCKET=0;II2_SINGLE_BUCKET=768;II2_SINGLE_BUCKET=0" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
