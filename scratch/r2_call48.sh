#!/bin/bash
# one-kernel point read: parity (+ extended fuzz), latency by mode
T=r05c
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -6 gpurun_out/${T}_tests.log
II2_FUZZ_SEEDS=300 II2_FUZZ_HEAVY_SEEDS=60 timeout 900 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_fuzz.log 2>&1; tail -4 gpurun_out/${T}_fuzz.log
timeout 900 python scratch/read_small.py --fracs 0.000001,0.00002 --env "II2_POINT_READ=2;II2_POINT_READ=1;II2_POINT_READ=0;II2_POINT_READ=2" > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
