#!/usr/bin/env python3
"""How much pinned host memory reaches the GPUs of one box when N ranks upload at once?
(VERDICT r01 "Next" 8: is 131 GB/s aggregate at N = 8 the box or the code?)
torchrun --nproc-per-node N scratch/h2d_nway.py   -> one JSON line on rank 0
Per rank: 1 GiB pinned -> HBM, (a) copy engine (cudaMemcpyAsync), (b) the library's own gather
path is SM-issued 16-byte loads over the bus, stood in for by a torch kernel reading the mapped
pinned buffer is not available from Python, so (b) = copy engine in 16 pieces (request overhead)."""
import json
import os
import time

import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", "0"))
world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 1 << 30
host = torch.empty(n, dtype=torch.uint8).pin_memory()
host.fill_(rank + 1)
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
back = torch.empty(n // 4, dtype=torch.uint8).pin_memory()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=5):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    barrier()
    return dt


def h2d():
    dev.copy_(host, non_blocking=True)


def h2d_pieces():
    step = n // 16
    for i in range(16):
        dev[i * step:(i + 1) * step].copy_(host[i * step:(i + 1) * step], non_blocking=True)


s2 = torch.cuda.Stream()


def duplex():  # upload 1 GiB while 256 MiB flows back (the shape of ii2_merge: 1.25 GB in, 0.3 out)
    dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        back.copy_(dev[: n // 4], non_blocking=True)
    s2.synchronize()


res = {}
for name, fn in (("h2d_1GiB", h2d), ("h2d_16_pieces", h2d_pieces), ("h2d_with_d2h_quarter", duplex)):
    dt = timed(fn)
    t = torch.tensor([n / dt / 1e9], device="cuda", dtype=torch.float64)
    if world > 1:
        allv = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allv, t)
        per = [float(x.item()) for x in allv]
    else:
        per = [float(t.item())]
    res[name] = {"per_rank_gbs": [round(x, 1) for x in per], "aggregate_gbs": round(sum(per), 1),
                 "min_gbs": round(min(per), 1)}
if rank == 0:
    numa = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node")] \
        if os.path.isdir("/sys/devices/system/node") else []
    print(json.dumps({"ranks": world, "cpus": os.cpu_count(), "numa_nodes": len(numa), "results": res}))
if world > 1:
    dist.destroy_process_group()
