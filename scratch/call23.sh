#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "intcomp or val or modes" 2>&1 | tail -3
bash scratch/call19.sh 2>&1 | sed -n 10,24p
