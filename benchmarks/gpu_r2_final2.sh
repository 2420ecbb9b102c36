#!/usr/bin/env bash
# round-2 closing visit: full GPU suite + smoke + bench line of the shipped defaults, then the two
# remaining N2/N3 candidates (selected through the environment) with their parity tests and A/B
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/final2_tests.log 2>&1; echo "suite rc=$?"; tail -4 gpurun_out/final2_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final2.json 2> gpurun_out/bench_final2.err; tail -2 gpurun_out/bench_final2.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final2.json'))
for k in ('value','ms_per_step','e2e','op_surface','grad_max_rel_err_vs_cpu','kernel_ms_per_step','cuda_graphs','loss_n2','density_n3','clocks','gpu_launches'):
    print(k, json.dumps(d.get(k))[:600])
PY
env GSB_DENSITY_VARIANT=2 GSB_LOSS_PGU=1 timeout 300 python -m pytest tests/test_loss.py tests/test_gpu_density.py -m gpu -q > gpurun_out/final2_exp_tests.log 2>&1; echo "experimental tests rc=$?"; tail -4 gpurun_out/final2_exp_tests.log
timeout 120 python benchmarks/ab_n2n3.py > gpurun_out/ab3_default.json 2>/dev/null; cat gpurun_out/ab3_default.json
env GSB_DENSITY_VARIANT=2 GSB_LOSS_PGU=1 timeout 120 python benchmarks/ab_n2n3.py > gpurun_out/ab3_exp.json 2>/dev/null; cat gpurun_out/ab3_exp.json
env GSB_DENSITY_VARIANT=2 timeout 120 ncu --set full --clock-control none --import-source on -k regex:"k_density_apply_rows" -c 1 -o gpurun_out/prof_n3v2_r2 python benchmarks/ab_n2n3.py --once > gpurun_out/ncu_n3v2.log 2>&1; tail -2 gpurun_out/ncu_n3v2.log
ls gpurun_out | head -5
