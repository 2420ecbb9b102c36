#!/usr/bin/env bash
# round-2 closing visit 2: the overflow fix (empty ranges after a capacity overflow) under the
# hardened tests, the full GPU suite, smoke, and the density rebuild with 6 / 12 loads in flight
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/final3_tests.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/final3_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
env GSB_DENSITY_VARIANT=3 timeout 200 python -m pytest tests/test_gpu_density.py -m gpu -q 2>&1 | tail -2
timeout 100 python benchmarks/ab_n2n3.py > gpurun_out/ab4_v2.json 2>/dev/null; cut -c1-900 gpurun_out/ab4_v2.json
env GSB_DENSITY_VARIANT=3 timeout 100 python benchmarks/ab_n2n3.py > gpurun_out/ab4_v3.json 2>/dev/null; cut -c1-900 gpurun_out/ab4_v3.json
