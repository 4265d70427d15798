#!/bin/bash
T=r02h
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q -m gpu > gpurun_out/${T}_mr_tests.log 2>&1; tail -15 gpurun_out/${T}_mr_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --strong-postings 400000000 --strong-cap 200000000 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err || tail -20 gpurun_out/${T}_bench_n2.err
python - <<'PY'
import json
b=json.load(open("gpurun_out/r02h_bench_n2.json"))
for k in ("value","ms_per_step","strong","cross_shard_read"):
    print(k, json.dumps(b.get(k))[:1500])
PY
