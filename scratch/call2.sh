#!/bin/bash
python -m inverted_index_2_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
bash scratch/sweep.sh "-DX_BASE" "-DK6_MIN_CTAS=8"
python -m inverted_index_2_b200.build --force > /dev/null 2>&1
python scratch/e2eprof.py 4 6 8 12 2>&1 | grep "^P"
echo NO_GATHER; II2_MERGE_NO_GATHER=1 python scratch/e2eprof.py 6 2>&1 | grep "^P"
