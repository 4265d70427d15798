#!/bin/bash
# round 2, call 1: parity of the fused bucket kernel + first timing, fused vs general kernels
T=r02a
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
timeout 600 python bench.py --verify --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/${T}_bench_fused.json 2> gpurun_out/${T}_bench_fused.err || tail -5 gpurun_out/${T}_bench_fused.err
II2_NO_FUSED=1 timeout 600 python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 > gpurun_out/${T}_bench_general.json 2> gpurun_out/${T}_bench_general.err || tail -5 gpurun_out/${T}_bench_general.err
cat gpurun_out/${T}_bench_fused.json | head -c 3000; echo; cat gpurun_out/${T}_bench_general.json | head -c 1500
