#!/bin/bash
# C4: which kernels of the intcomp codec take the time (launch list under ncu)
T=r03f
timeout 600 python bench_extra.py --which c4 --c4-values 16777216 > gpurun_out/${T}_c4_plain.json 2> gpurun_out/${T}_c4.err || tail -5 gpurun_out/${T}_c4.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${T}_c4_launches.csv \
  python bench_extra.py --which c4 --c4-values 16777216 > gpurun_out/${T}_ncu.log 2>&1 || tail -3 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_*
