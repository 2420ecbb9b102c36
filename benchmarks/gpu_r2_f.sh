#!/usr/bin/env bash
# 8-GPU A/B: colour all-gather started early (under the per-Gaussian backward) vs in reduce()
set -x
mkdir -p gpurun_out
for eg in 0 1 0 1; do
GSB_EARLY_GATHER=$eg timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_8gpu_eg$eg.json 2> gpurun_out/bench_8gpu_eg$eg.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_8gpu_eg$eg.json'))
print('early_gather=$eg', 'value', d['value'], 'ms', d['ms_per_step'], 'config5', d['config5']['ms_per_step'])
PY
done
