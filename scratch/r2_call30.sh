#!/bin/bash
# warp-per-term kernel: occupancy 4/5/6 CTAs per SM, 64 values per source in flight
T=r03a
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
L=scratch/variants/libii2_mw4.so,scratch/variants/libii2_mw5.so,scratch/variants/libii2_mw6.so
for ML in 64 512; do
timeout 600 python scratch/sweep2.py --libs $L --terms 200000 --segments 64 --postings 200000000 --steps 3 --max-len $ML > gpurun_out/${T}_dense_ml$ML.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense_ml$ML.jsonl
done
