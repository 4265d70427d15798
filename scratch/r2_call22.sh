#!/bin/bash
T=r02u
V=scratch/variants
timeout 900 python scratch/sweep2.py --libs default:768,$V/libii2_k2b10.so:768,$V/libii2_k2b12.so:768,$V/libii2_k2b6.so:768,default:768 --steps 10 > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
