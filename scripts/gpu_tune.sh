#!/bin/bash
# dev: time library variants of the element-stiffness / replay kernels.  usage: gpu_tune.sh n variant...
n=$1; shift
mkdir -p gpurun_out
for v in default "$@"; do
  if [ "$v" = default ]; then unset EASYFEA_B200_LIB; else export EASYFEA_B200_LIB=$PWD/build/variants/$v.so; fi
  python scripts/tune_ke.py HEXA8 $n 10 2>&1 | tail -1 | tee -a gpurun_out/tune.log
done
