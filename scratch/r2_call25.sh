#!/bin/bash
T=r02x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/${T}_mr_tests.log 2>&1; tail -5 gpurun_out/${T}_mr_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "comm or medium or val_views" > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
