set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; tail -3 gpurun_out/bench1.err; cat gpurun_out/bench1.json
