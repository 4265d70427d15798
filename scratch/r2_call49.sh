#!/bin/bash
T=r05d
for ML in 64 512; do
timeout 600 python scratch/sweep2.py --terms 200000 --segments 64 --postings 200000000 --steps 3 --max-len $ML > gpurun_out/${T}_dense_ml$ML.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense_ml$ML.jsonl
done
timeout 600 python scratch/sweep2.py --terms 100000 --segments 64 --postings 250000000 --steps 3 --max-len 512 > gpurun_out/${T}_dense_2500.jsonl 2> gpurun_out/${T}_dense.err || tail -5 gpurun_out/${T}_dense.err
cat gpurun_out/${T}_dense_2500.jsonl
