#!/bin/bash
# 8 GPUs: N-way H2D ceiling, bench at N = 8 (weak + strong 1B + cross-shard read through the C-ABI)
T=r02k
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n scratch/h2d_nway.py >> gpurun_out/${T}_h2d_nway.jsonl 2>> gpurun_out/${T}_h2d_nway.err
done
cat gpurun_out/${T}_h2d_nway.jsonl
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err || tail -20 gpurun_out/${T}_bench_n8.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02k_bench_n8.json"))
for k in ("value","ms_per_step","e2e","strong","cross_shard_read"):
    print(k, json.dumps(b.get(k))[:1600])
PY
