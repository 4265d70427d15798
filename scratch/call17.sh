#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
II2_MERGE_TAPER=0 timeout 200 python scratch/e2eprof.py 6 2>&1 | tail -1
timeout 200 python scratch/e2eprof.py 6 7 8 2>&1 | tail -3
