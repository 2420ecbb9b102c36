#!/usr/bin/env bash
# round-2 final 1-GPU visit: full GPU suite, smoke, bench line (both arms), ncu evidence of the shipped binary
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_arm.json 2> gpurun_out/bench_ref_arm.err; cat gpurun_out/bench_ref_arm.json | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_1gpu.json'))
for k in ('value','ms_per_step','e2e','op_surface','grad_max_rel_err_vs_cpu','roofline','kernel_ms_per_step','cuda_graphs','ref_gpu','config5','clocks','gpu_launches','loss_n2'):
    print(k, json.dumps(d.get(k))[:500])
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 320 --csv --log-file gpurun_out/launches.csv python benchmarks/profile_step.py 30 fused > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_draw3|k_draw_bwd4" -s 2 -c 2 -o gpurun_out/prof_draw_r2 python benchmarks/profile_step.py 3 fused > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none -k regex:"k_radix_pass|k_colscan|k_keys|k_rects_scan|k_tile_list|k_ranges|k_preprocess" -s 14 -c 14 -o gpurun_out/prof_small_r2 python benchmarks/profile_step.py 3 fused > gpurun_out/ncu_small.log 2>&1
timeout 300 python benchmarks/run_reference_scripts.py --impl ours > gpurun_out/ref_scripts.log 2>&1; grep -c "OK" gpurun_out/ref_scripts.log
ls gpurun_out | head -40
