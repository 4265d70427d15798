#!/bin/bash
# 8 GPUs, final: bench at N = 8 (weak + strong 1B + cross-shard read) with the end-of-round kernels
T=r04v
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err || tail -20 gpurun_out/${T}_bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err || tail -20 gpurun_out/${T}_bench_n2.err
python - <<'PY'
import json
for n in (8,2):
    b=json.load(open("gpurun_out/r04v_bench_n%d.json"%n))
    print("N",n,"value",b["value"],"ms",b["ms_per_step"],"e2e",b["e2e"]["value"],b["e2e"]["ms_per_step"])
    print("  strong", json.dumps(b["strong"]["merge"]), json.dumps(b["strong"]["build"]), b["strong"]["imbalance_max_over_mean"], b["strong"]["total_postings"])
    print("  xread", b["cross_shard_read"]["us_per_read"], b["cross_shard_read"]["verified"], b["cross_shard_read"]["gathered_postings"])
PY
