#!/usr/bin/env bash
# round-2 visit: graphed step test + config-4 sweep with CUDA graphs
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -15
timeout 900 python benchmarks/sweep_config4.py --iters 10 > gpurun_out/sweep.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/sweep_config4.json'))['rows']:
    k=r['kernel_ms']; print(r['N'], r['W'], 'ms/step %.3f graphs %s draw %.3f (%.0f%%) drawB %.3f (%.0f%%) sort %.3f'%(r['ms_per_step'],r['ms_per_step_cuda_graphs'],k['draw'],100*r['draw_hbm_frac'],k['draw_backward'],100*r['draw_backward_hbm_frac'],k['sort']))
PY
