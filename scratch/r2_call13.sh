#!/bin/bash
T=r02o
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 600 python scratch/sweep2.py --terms 200000 --segments 64 --postings 200000000 --steps 3 > gpurun_out/${T}_dense.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense.jsonl
timeout 600 python scratch/sweep2.py --terms 400000 --segments 64 --postings 200000000 --steps 3 > gpurun_out/${T}_dense2.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense2.jsonl
