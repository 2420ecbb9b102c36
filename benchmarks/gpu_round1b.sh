set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 900 python benchmarks/compare_ref_gpu.py --iters 5 > gpurun_out/compare.log 2>&1; tail -60 gpurun_out/compare.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python benchmarks/profile_step.py 2 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_draw -s 2 -c 2 -o gpurun_out/prof_draw_r1 python benchmarks/profile_step.py 2 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
