#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for w in 1 2 4 1000; do echo "II2_K1B_WAVES=$w"; II2_K1B_WAVES=$w bash scratch/sweep.sh "-DX_BASE"; done
