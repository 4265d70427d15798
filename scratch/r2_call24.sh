#!/bin/bash
T=r02w
timeout 900 python -m pytest tests/test_gpu_fuzz.py -q -m gpu > gpurun_out/${T}_fuzz.log 2>&1; tail -15 gpurun_out/${T}_fuzz.log
