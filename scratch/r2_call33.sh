#!/bin/bash
# bound for VERDICT "Next" 2: K2b with the sort removed (timing only, results differ)
T=r03e
timeout 600 python scratch/sweep2.py --libs default,scratch/variants/libii2_nosort.so,scratch/variants/libii2_nolocal.so --steps 6 > gpurun_out/${T}_k2b_nosort.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_k2b_nosort.jsonl
