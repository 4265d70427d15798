#!/bin/bash
# one compute-sanitizer tool per call ($1 = memcheck | racecheck)
T=r02_sanitizer_$1
timeout 300 python scratch/sanitize_run.py > gpurun_out/${T}_plain.log 2>&1 || { tail -20 gpurun_out/${T}_plain.log; exit 1; }
timeout 2400 compute-sanitizer --tool $1 --log-file gpurun_out/${T}.log python scratch/sanitize_run.py > gpurun_out/${T}_run.log 2>&1
echo "exit $?"; tail -5 gpurun_out/${T}_run.log; tail -30 gpurun_out/${T}.log
