#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python -c "
import json; d=json.load(open('gpurun_out/bench.json')); print('value', d['value'], 'op_surface', d['op_surface'], 'e2e', d['e2e']['value'])"
timeout 1200 python benchmarks/compare_ref_gpu.py --iters 10 --configs config2 > gpurun_out/compare.log 2>&1; grep -E "t_fwd_ms|t_bwd_ms|mpix_per_s|speedup|config2\"" -A1 gpurun_out/compare.log | head -40
