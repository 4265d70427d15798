#!/bin/bash
# round-1 late experiments: K6 batched copy, K1b prefetch / descriptor staging, bucket target
python -m inverted_index_2_b200.build --force > /dev/null 2>&1 || echo BUILD FAILED
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash scratch/sweep.sh "-DX_BASE" "-DK6_SERIAL_COPY" "-DK6_MIN_CTAS=6" "-DK1B_PREFETCH_POST" "-DK1B_SEG_SMEM" "-DK1B_SEG_SMEM -DK1B_PREFETCH_POST"
for b in 640 704; do echo "II2_BUCKET=$b"; II2_BUCKET=$b bash scratch/sweep.sh "-DX_BASE"; done
python - <<'PY'
import torch, time
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
e0.record()
for _ in range(5): h.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H pinned GB/s", 5 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9)
PY
