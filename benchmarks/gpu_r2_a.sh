#!/usr/bin/env bash
# round-2 visit: backward variants: parity tests, A/B, ncu capture of the default
set -x
mkdir -p gpurun_out
./benchmarks/micro/rcp_check
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py -m gpu -x -q 2>&1 | tail -8
GSB_AB_ENVS="v2:GSB_BWD_VARIANT=2" timeout 900 python benchmarks/ab_variants.py 2>&1 | tail -12
cp gpurun_out/ab_variants.json gpurun_out/ab_r2_d.json
ncu --set full --clock-control none --import-source on -k regex:k_draw_bwd -s 1 -c 1 -o gpurun_out/prof_bwd4c python benchmarks/profile_step.py 2 fused > gpurun_out/ncu_bwd4.log 2>&1
