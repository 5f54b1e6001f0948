#!/bin/bash
# dev: A/B the homogeneous-C stiffness kernel forms on the GPU box.  usage: gpu_ab.sh n form...
n=${1:-128}; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_operators.py -x -q 2>&1 | tail -3
for f in "$@"; do
  EFB_ELASTIC_KERNEL=$f TUNE_REPLAY=${TUNE_REPLAY:-0} python scripts/tune_ke.py HEXA8 $n 10 2>&1 | tail -1 | sed "s/^/$f /" | tee -a gpurun_out/ab.log
done
