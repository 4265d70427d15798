#!/bin/bash
# round 2, call 2: parity of the slot-scheduled union phase, bucket-size + variant sweep, ncu of k12f
T=r02b
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
timeout 600 python scratch/sweep2.py --libs ,scratch/variants/libii2_t256.so,scratch/variants/libii2_t512c1.so --buckets 448,512,576,640,704,768 > gpurun_out/${T}_sweep.jsonl 2> gpurun_out/${T}_sweep.err || tail -5 gpurun_out/${T}_sweep.err
cat gpurun_out/${T}_sweep.jsonl
II2_BUCKET=576 timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_plain.log 2>&1 && \
II2_BUCKET=576 timeout 900 ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k12f_bucket_kernel|k6_dense_kernel|k1_partition_chunks_raw|k1_bucket_stats' \
  --launch-skip 4 --launch-count 4 -f -o gpurun_out/${T}_full \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${T}_ncu.log 2>&1 || tail -5 gpurun_out/${T}_ncu.log
ls -la gpurun_out/${T}_*
