#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -3
timeout 400 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu -k "c4 or C4 or codec or bitmask" 2>&1 | tail -3
bash scratch/call19.sh 2>&1 | head -18
