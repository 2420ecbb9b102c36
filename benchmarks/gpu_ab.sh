#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
GSB_RASTER_VARIANT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
GSB_RASTER_VARIANT=1 timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_variant1.json 2>/dev/null; cat gpurun_out/bench_variant1.json | python -c "import sys,json; d=json.load(sys.stdin); print('variant1', d['value'], d['kernel_ms_per_step'])"
