#!/bin/bash
# N-GPU bench line (run with gpurun --gpus N): python -m torch.distributed.run as the driver launches it
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; tail -5 gpurun_out/bench_n$N.err
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/pytest_multi_r2.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_multi_r2.log
