#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "pipelined" 2>&1 | tail -3
timeout 400 python scratch/e2emodes.py - dma dma2 dma3 dma4 hybrid1 hybrid2 hybrid3 -::128 -::32 -:4 -:8 dma2:4 dma2:8 dma3:4 dma4:3 hybrid2:4 hybrid2:8 hybrid2:6:32 - 2>&1 | grep -v "^$" | tail -25 | tee gpurun_out/e2emodes.log
