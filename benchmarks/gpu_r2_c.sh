#!/usr/bin/env bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k reference_backward 2>&1 | tail -40
timeout 900 python benchmarks/compare_ref_gpu.py --three-way --out gpurun_out/compare_ref_gpu > gpurun_out/three_way.log 2>&1; tail -3 gpurun_out/three_way.log
