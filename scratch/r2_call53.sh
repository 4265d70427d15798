#!/bin/bash
T=r05g
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "point or read_range or window or many_short" > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 900 python scratch/read_small.py --fracs 0.000001,0.00002,0.001 > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
