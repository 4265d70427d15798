#!/bin/bash
bash scratch/sweep.sh "-DK1B_RSTART_W0 -DK1B_RANK_SUB"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
