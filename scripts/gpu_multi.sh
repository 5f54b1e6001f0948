#!/bin/bash
# Run on an N-GPU box (gpurun --gpus N): the NCCL parity test and the torchrun bench line.  usage: gpu_multi.sh <tag> <N> [bench args]
tag=${1:-m2}; N=${2:-2}; shift 2 || true
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/${tag}_smi.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log; tail -4 gpurun_out/${tag}_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
  bench.py --gpus $N "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
echo "bench exit $?"; tail -c 2500 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err
