#!/bin/bash
# dev: A/B replay forms.  usage: gpu_ab3.sh n form...
n=${1:-160}; shift
mkdir -p gpurun_out
python -m pytest tests/test_gpu_assembly.py -x -q 2>&1 | tail -3
for f in "$@"; do
  EFB_REPLAY_KERNEL=$f TUNE_REPLAY=1 python scripts/tune_ke.py HEXA8 $n 10 2>&1 | tail -1 | sed "s/^/$f /" | tee -a gpurun_out/ab3.log
  EFB_REPLAY_KERNEL=$f TUNE_REPLAY=1 python scripts/tune_ke.py TETRA4 100 10 2>&1 | tail -1 | sed "s/^/$f /" | tee -a gpurun_out/ab3.log
done
