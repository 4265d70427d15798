#!/bin/bash
python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
python scratch/c3prof.py 2>&1 | tail -6 | grep unprofiled
python bench_extra.py --which prefix 2>/dev/null | python -c "
import json,sys
for ln in sys.stdin:
    d=json.loads(ln)
    for r in d['results']: print(r['prefix_len'], r['prefixes'], round(r['median_us'],1), r['phase_us'])
"
