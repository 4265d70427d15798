#!/bin/bash
# end-of-round evidence: tests, bench lines, ncu launch list + --set full captures (tag = $1)
T=${1:-r01d}
python -m pytest tests -x -q -m gpu > gpurun_out/${T}_tests.log 2>&1; tail -2 gpurun_out/${T}_tests.log
python bench.py --verify > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err || tail -3 gpurun_out/${T}_bench.err
python bench.py --impl reference > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_ref.err || tail -3 gpurun_out/${T}_ref.err
python bench_extra.py --which c1,c3,c4,prefix > gpurun_out/${T}_bench_extra.json 2> gpurun_out/${T}_extra.err || tail -3 gpurun_out/${T}_extra.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1 || tail -3 gpurun_out/ncu_l.log
if [ "$2" = "full" ]; then
ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k1b_group_kernel|k2b_union_kernel|k6_emit_kernel|k1_partition_chunks_raw|k1_bucket_stats' \
  --launch-skip 5 --launch-count 5 -f -o gpurun_out/${T}_full \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_f.log 2>&1 || tail -3 gpurun_out/ncu_f.log
fi
python scratch/c4_one.py > gpurun_out/c4_one.log 2>&1 && \
ncu --set full --clock-control none --import-source on \
  --kernel-name regex:'k_dec_walk|k_dec_blocks|k_dec_blocksum|k_enc_emit_huge|k_enc_size_huge|k_gather_host' \
  --launch-count 12 -f -o gpurun_out/${T}_codec \
  python scratch/c4_one.py > gpurun_out/ncu_c.log 2>&1 || tail -3 gpurun_out/ncu_c.log
ls -la gpurun_out/${T}_*
