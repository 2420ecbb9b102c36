#!/usr/bin/env bash
# round-2 visit (N GPUs): bench line at N = $1
set -x
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; tail -3 gpurun_out/bench_${N}gpu.err; python - <<PY
import json
d=json.load(open('gpurun_out/bench_${N}gpu.json'))
for k in ('value','ms_per_step','e2e','allreduce','config5'):
    print(k, json.dumps(d.get(k))[:1200])
PY
