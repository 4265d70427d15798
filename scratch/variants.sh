#!/bin/bash
# Build several variants of libii2.so HERE (no GPU needed), then time them in ONE gpurun call.
#   local:   bash scratch/variants.sh build "<tag>=<nvcc flags>" ...     -> scratch/variants/libii2_<tag>.so
#   GPU box: bash scratch/variants.sh run [bench.py flags]               -> one line per variant
# Costs on this pool (measured in round 1): a bench-only call ~45 s of the GPU budget, the parity
# file tests/test_gpu_parity.py ~25 s, the whole `pytest -m gpu` suite ~3.3 min (full-size cases).
set -e
cd "$(dirname "$0")/.."
mode=$1; shift
mkdir -p scratch/variants gpurun_out
if [ "$mode" = build ]; then
  for spec in "$@"; do
    tag=${spec%%=*}; flags=${spec#*=}
    II2_NVCC_EXTRA="$flags" python -m inverted_index_2_b200.build > /dev/null
    cp inverted_index_2_b200/libii2.so scratch/variants/libii2_${tag}.so
    echo "built $tag ($flags)"
  done
  python -m inverted_index_2_b200.build --force > /dev/null 2>&1  # leave the DEFAULT build in place (forced: the objects are newer than the sources)
else
  cp inverted_index_2_b200/libii2.so scratch/variants/.default.so
  for so in scratch/variants/libii2_*.so; do
    tag=$(basename $so .so); tag=${tag#libii2_}
    cp $so inverted_index_2_b200/libii2.so
    python bench.py --no-cpu-baseline --no-e2e --steps 10 --warmup 3 "$@" > gpurun_out/v.json 2> gpurun_out/v.err \
      || { echo "$tag RUN FAILED"; tail -2 gpurun_out/v.err; continue; }
    python - "$tag" <<'PY'
import json, sys
b = json.load(open("gpurun_out/v.json"))
print(sys.argv[1], "| ms", round(b["ms_per_step"], 3),
      " ".join("%s=%.3f" % (k["name"][:9], k["ms"] / k["count"]) for k in b["kernels"][1:]))
PY
  done
  cp scratch/variants/.default.so inverted_index_2_b200/libii2.so
fi
