#!/bin/bash
# windows kernel with two warps per segment; whole gpu suite after the launch-chain changes
T=r03l
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -4 gpurun_out/${T}_tests.log
timeout 900 python scratch/read_small.py --fracs 0.000001,0.00002,0.001,0.01,0.1 > gpurun_out/${T}_reads.jsonl 2> gpurun_out/${T}.err || tail -5 gpurun_out/${T}.err
cat gpurun_out/${T}_reads.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
