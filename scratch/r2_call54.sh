#!/bin/bash
T=r05h
for i in 1 2 3; do timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "concurrent" > gpurun_out/${T}_tests$i.log 2>&1; tail -2 gpurun_out/${T}_tests$i.log; done
