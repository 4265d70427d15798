#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for sb in 0 128 192 320; do echo "II2_SMALL_BUCKET=$sb"; II2_SMALL_BUCKET=$sb python scratch/c3prof.py 2>&1 | tail -6; done
