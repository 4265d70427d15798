#!/bin/bash
T=r05e
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "many_short or point" > gpurun_out/${T}_tests.log 2>&1; tail -6 gpurun_out/${T}_tests.log
