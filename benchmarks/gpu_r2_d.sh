#!/usr/bin/env bash
# round-2 visit (2 GPUs): 2-GPU bench line
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -3 gpurun_out/bench_2gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
for k in ('value','ms_per_step','e2e','allreduce','config5','kernel_ms_per_step'):
    print(k, json.dumps(d.get(k))[:1100])
PY
