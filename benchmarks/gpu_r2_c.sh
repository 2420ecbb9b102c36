#!/usr/bin/env bash
# round-2 visit: GPU suite, sparse corner after the cheaper tile list, ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python benchmarks/sweep_config4.py --iters 10 > gpurun_out/sweep.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/sweep_config4.json'))['rows']:
    k=r['kernel_ms']; print(r['N'], r['W'], 'ms/step %.3f draw %.3f (%.0f%%) drawB %.3f (%.0f%%) sort %.3f'%(r['ms_per_step'],k['draw'],100*r['draw_hbm_frac'],k['draw_backward'],100*r['draw_backward_hbm_frac'],k['sort']))
PY
python benchmarks/ab_variants.py 2>&1 | tail -3
