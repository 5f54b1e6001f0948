#!/bin/bash
# dev: A/B the replay forms (+ the default stiffness kernel) and run the assembly parity tests.  usage: gpu_ab2.sh n
n=${1:-160}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_assembly.py tests/test_gpu_operators.py -x -q 2>&1 | tail -3
for f in fast ring; do
  EFB_REPLAY_KERNEL=$f TUNE_REPLAY=1 python scripts/tune_ke.py HEXA8 $n 10 2>&1 | tail -1 | sed "s/^/$f /" | tee -a gpurun_out/ab2.log
done
for e in TETRA4 TRI3 HEXA27; do
  nn=$([ $e = HEXA27 ] && echo 40 || ([ $e = TRI3 ] && echo 1000 || echo 100))
  for f in fast ring; do
  EFB_REPLAY_KERNEL=$f TUNE_REPLAY=1 python scripts/tune_ke.py $e $nn 10 2>&1 | tail -1 | sed "s/^/$f /" | tee -a gpurun_out/ab2.log
  done
done
