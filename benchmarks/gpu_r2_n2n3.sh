#!/usr/bin/env bash
# round-2 N2/N3 visit: parity of the row-streaming loss kernels, the row-group density rebuild and
# the hand-written slots scan (selected through the environment), A/B timings, ncu of the new kernels
set -x
mkdir -p gpurun_out
NEW="GSB_LOSS_VARIANT=1 GSB_DENSITY_VARIANT=1 GSB_DENSITY_SCAN_VARIANT=1"
env $NEW timeout 400 python -m pytest tests/test_loss.py tests/test_gpu_density.py -m gpu -x -q > gpurun_out/n2n3_tests.log 2>&1; echo "new-variant tests rc=$?"; tail -5 gpurun_out/n2n3_tests.log
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -x -q -k "record_cache or fused_shortcuts or graphed or capacity" 2>&1 | tail -3
env GSB_LOSS_VARIANT=0 GSB_DENSITY_VARIANT=0 GSB_DENSITY_SCAN_VARIANT=0 timeout 200 python benchmarks/ab_n2n3.py > gpurun_out/ab_n2n3_old.json 2> gpurun_out/ab_old.err; cat gpurun_out/ab_n2n3_old.json; tail -2 gpurun_out/ab_old.err
env $NEW timeout 200 python benchmarks/ab_n2n3.py > gpurun_out/ab_n2n3_new.json 2> gpurun_out/ab_new.err; cat gpurun_out/ab_n2n3_new.json; tail -2 gpurun_out/ab_new.err
for sh in 42 168; do env GSB_LOSS_VARIANT=1 GSB_LOSS_STRIP=$sh timeout 100 python benchmarks/ab_n2n3.py --skip-density > gpurun_out/ab_loss_strip$sh.json 2>/dev/null; cut -c1-600 gpurun_out/ab_loss_strip$sh.json; done
env GSB_LOSS_VARIANT=1 timeout 100 python benchmarks/ab_n2n3.py --skip-density --hw 2160x3840 > gpurun_out/ab_loss_4k_new.json 2>/dev/null; cut -c1-600 gpurun_out/ab_loss_4k_new.json
env $NEW timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ssim_fwd_rows|k_ssim_bwd_rows|k_density_apply_rows|k_density_slots|k_density_classify_count|k_density_scan_blocks" -c 6 -o gpurun_out/prof_n2n3_r2 python benchmarks/ab_n2n3.py --once > gpurun_out/ncu_n2n3.log 2>&1; tail -3 gpurun_out/ncu_n2n3.log
ls -la gpurun_out | head -30
