#!/usr/bin/env bash
# One GPU-box visit: build check, GPU parity tests, smoke, bench line, optional extras.
#   bash benchmarks/gpu_check.sh [tests] [bench] [compare] [ncu] [sanitize]
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for what in "$@"; do
  case $what in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
      python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ;;
    bench)
      timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
      tail -5 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
    compare)
      timeout 1200 python benchmarks/compare_ref_gpu.py --iters 10 > gpurun_out/compare.log 2>&1
      tail -80 gpurun_out/compare.log ;;
    ncu)
      ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv \
          python benchmarks/profile_step.py 2 fused > gpurun_out/ncu_launch.log 2>&1
      ncu --set full --clock-control none --import-source on -k regex:k_draw -s 2 -c 2 -o gpurun_out/prof_draw \
          python benchmarks/profile_step.py 2 fused > gpurun_out/ncu_full.log 2>&1
      tail -3 gpurun_out/ncu_full.log ;;
    sanitize)
      timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py tests/test_gpu_fused.py \
          tests/test_gpu_exchange.py tests/test_gpu_density.py tests/test_loss.py -m gpu -x -q \
          -k "not config2 and not 50000 and not 100000 and not full_size and not 70001 and not 1080" \
          > gpurun_out/sanitize_memcheck.log 2>&1
      grep -E "=========" gpurun_out/sanitize_memcheck.log | grep -v "Host Frame" | head -30
      tail -3 gpurun_out/sanitize_memcheck.log
      timeout 900 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
          -k "splat_and_splatB_vs_oracle and 200" > gpurun_out/sanitize_racecheck.log 2>&1
      grep -E "=========" gpurun_out/sanitize_racecheck.log | grep -v "Host Frame" | head -20
      tail -3 gpurun_out/sanitize_racecheck.log ;;
  esac
done
ls -la gpurun_out
