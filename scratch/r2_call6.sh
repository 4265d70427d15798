#!/bin/bash
T=r02f
export II2_COALESCE=1
V=scratch/variants
timeout 900 python scratch/sweep2.py --libs default:768,$V/libii2_c512t256m4.so:352+384+416,$V/libii2_c512t128m4.so:352+384+416,$V/libii2_c384t128m5.so:256+288+320,$V/libii2_c384t256m5.so:256+288+320,$V/libii2_c768t128m2.so:512+576 > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
