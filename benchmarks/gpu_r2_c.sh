#!/usr/bin/env bash
# round-2 visit: GPU suite, config-4 sweep after the tile-list pass, ncu launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python benchmarks/sweep_config4.py --iters 10 > gpurun_out/sweep.log 2>&1; python - <<'PY'
import json
for r in json.load(open('gpurun_out/sweep_config4.json')):
    k=r['kernel_ms']; print(r['N'], r['W'], 'ms/step %.3f draw %.3f (%.0f%%) drawB %.3f (%.0f%%)'%(r['ms_per_step'],k['draw'],100*r['draw_hbm_frac'],k['draw_backward'],100*r['draw_backward_hbm_frac']))
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 300 --csv --log-file gpurun_out/launches.csv python benchmarks/profile_step.py 30 fused > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches.csv
