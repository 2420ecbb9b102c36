#!/usr/bin/env bash
# extras: config-4 sweep, training stand-in, reference comparison
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" 2>&1 | tail -2
for what in "$@"; do
  case $what in
    sweep) timeout 1500 python benchmarks/sweep_config4.py --iters 8 > gpurun_out/sweep.log 2>&1; tail -12 gpurun_out/sweep.log | cut -c1-400 ;;
    train) timeout 900 python benchmarks/train_synthetic.py --n 100000 --views 4 --iters 200 --size 640x360 > gpurun_out/train.log 2>&1; tail -14 gpurun_out/train.log ;;
    traintest) timeout 900 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -5 ;;
    n3tests) timeout 900 python -m pytest tests/test_gpu_density.py tests/test_gpu_parity.py -m gpu -x -q -k "density or gau_io or degenerate or plan_and_apply or full_size or controller or huge" 2>&1 | tail -25 ;;
    exchange) timeout 600 python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -25 ;;
    density) timeout 900 python benchmarks/compare_density_ref.py --n 1000000 > gpurun_out/density.log 2>&1; tail -60 gpurun_out/density.log ;;
    trainref) timeout 2400 python benchmarks/train_reference.py --n 20000 --views 8 --size 320x240 > gpurun_out/trainref.log 2>&1; tail -30 gpurun_out/trainref.log ;;
    mgpu) N=$(nvidia-smi -L | wc -l); timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_gpus$N.json 2> gpurun_out/bench_gpus$N.err; tail -5 gpurun_out/bench_gpus$N.err; cat gpurun_out/bench_gpus$N.json ;;
  esac
done
