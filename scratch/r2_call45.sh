#!/bin/bash
T=r04z
II2_FUZZ_SEEDS=1 II2_FUZZ_HEAVY_SEEDS=150 timeout 1200 python -m pytest tests/test_gpu_fuzz.py -x -q -m gpu -k heavy > gpurun_out/${T}_fuzz_heavy150.log 2>&1; tail -8 gpurun_out/${T}_fuzz_heavy150.log
