#!/usr/bin/env bash
# round-2 visit: forward variant 3: full GPU suite, A/B timing vs variant 2
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
GSB_AB_ENVS="fwd2:GSB_FWD_VARIANT=2" timeout 600 python benchmarks/ab_variants.py 2>&1 | tail -6
python - <<'PY'
import json; d=json.load(open('gpurun_out/ab_variants.json'))
for k,v in d.items(): print(k, json.dumps(v.get('kernels')))
PY
