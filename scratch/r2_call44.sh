#!/bin/bash
T=r04z
II2_FUZZ_SEEDS=1500 timeout 1200 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_fuzz1500.log 2>&1; tail -5 gpurun_out/${T}_fuzz1500.log
