#!/bin/bash
bash scratch/sweep.sh "-DX_BASE" "-DK1B_FUSED_HEAD"
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
