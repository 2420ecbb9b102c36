#!/usr/bin/env bash
# round-2 N2/N3 visit 2: pipelined (two rows in flight) loss kernels, coalesced row-group density
# rebuild; parity, A/B against the previous kernels on the same box, ncu of the new kernels
set -x
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_loss.py tests/test_gpu_density.py -m gpu -x -q > gpurun_out/n2n3b_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/n2n3b_tests.log
timeout 200 python benchmarks/ab_n2n3.py > gpurun_out/ab2_new.json 2> gpurun_out/ab2_new.err; cat gpurun_out/ab2_new.json; tail -2 gpurun_out/ab2_new.err
env GSB_LOSS_VARIANT=0 GSB_DENSITY_VARIANT=0 GSB_DENSITY_SCAN_VARIANT=0 timeout 200 python benchmarks/ab_n2n3.py > gpurun_out/ab2_old.json 2>/dev/null; cat gpurun_out/ab2_old.json
for v in lossb5 lossb6; do env GSB_LIB=$PWD/easygaussiansplatting_b200/libgsplat_b200_$v.so timeout 100 python benchmarks/ab_n2n3.py --skip-density > gpurun_out/ab2_$v.json 2>/dev/null; cut -c1-500 gpurun_out/ab2_$v.json; done
for sh in 56 120; do env GSB_LOSS_STRIP=$sh timeout 100 python benchmarks/ab_n2n3.py --skip-density > gpurun_out/ab2_strip$sh.json 2>/dev/null; cut -c1-500 gpurun_out/ab2_strip$sh.json; done
timeout 100 python benchmarks/ab_n2n3.py --skip-density --hw 2160x3840 > gpurun_out/ab2_4k.json 2>/dev/null; cut -c1-500 gpurun_out/ab2_4k.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_density_apply_rows|k_density_slots|k_density_classify_count|k_density_scan_blocks" -c 4 -o gpurun_out/prof_n3_r2 python benchmarks/ab_n2n3.py --once > gpurun_out/ncu_n3.log 2>&1; tail -2 gpurun_out/ncu_n3.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"k_ssim_fwd_rows|k_ssim_bwd_rows" -c 2 -o gpurun_out/prof_n2_r2 python benchmarks/ab_n2n3.py --once --skip-density > gpurun_out/ncu_n2.log 2>&1; tail -2 gpurun_out/ncu_n2.log
ls -la gpurun_out | head -30
