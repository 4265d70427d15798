#!/bin/bash
T=r02d
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 900 python scratch/sweep2.py --libs default:704+768+832+896+960 --env "II2_COALESCE=4;II2_COALESCE=1;II2_COALESCE=8" > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
