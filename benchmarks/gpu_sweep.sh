#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 1500 python benchmarks/sweep_config4.py --iters 8 > gpurun_out/sweep.log 2>&1; tail -3 gpurun_out/sweep.log | cut -c1-300
