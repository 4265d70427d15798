#!/bin/bash
# speculative point reads: parity, latency
T=r03o
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -6 gpurun_out/${T}_tests.log
timeout 900 python scratch/read_small.py --fracs 0.000001,0.00002,0.001 --env "II2_POINT_READ=1;II2_POINT_READ=0;II2_POINT_READ=1;II2_POINT_READ=0" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
