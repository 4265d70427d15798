#!/bin/bash
T=r01d
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
python bench_extra.py --which c4 > gpurun_out/${T}_bench_extra_c4.json 2> gpurun_out/${T}_extra.err || tail -3 gpurun_out/${T}_extra.err
ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k_dec_tile|k_dec_blocks|k_dec_blocksum|k_enc_emit_huge|k_enc_size_huge' \
  --launch-count 9 -f -o gpurun_out/${T}_codec \
  python scratch/c4_one.py > gpurun_out/ncu_c.log 2>&1 || tail -3 gpurun_out/ncu_c.log
ls -la gpurun_out/${T}_*
