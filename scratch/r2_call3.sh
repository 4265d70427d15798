#!/bin/bash
T=r02c
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
V=scratch/variants
timeout 900 python scratch/sweep2.py --libs default:704+768,$V/libii2_c1024t256.so:704+768,$V/libii2_c768t256m3.so:512+576,$V/libii2_c640t256m3.so:416+448+480,$V/libii2_c640t384m3.so:416+448+480 > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
