#!/bin/bash
T=r05x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/${T}_mr_tests.log 2>&1; tail -5 gpurun_out/${T}_mr_tests.log
