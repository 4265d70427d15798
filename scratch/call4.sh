#!/bin/bash
bash scratch/sweep.sh "-DK1B_RANK_THREAD" "-DK1B_RSTART_W0" "-DK1B_RANK_THREAD -DK1B_RSTART_W0"
II2_NVCC_EXTRA="-DK1B_RANK_THREAD -DK1B_RSTART_W0" python -m inverted_index_2_b200.build > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
