#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python scratch/c3prof.py 2>&1 | tail -6
