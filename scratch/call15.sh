#!/bin/bash
bash scratch/sweep.sh "-DK2B_STAGE"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
