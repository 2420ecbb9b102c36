#!/usr/bin/env bash
# round-2 visit: hand-written binning (scan + radix sort): parity tests, per-kernel timing
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python benchmarks/ab_variants.py 2>&1 | tail -6
python - <<'PY'
import json; d=json.load(open('gpurun_out/ab_variants.json'))
for k,v in d.items(): print(k, json.dumps(v.get('kernels')))
PY
