#!/bin/bash
# dev: ncu --set full of the homogeneous-C stiffness kernel forms.  usage: gpu_ncu_ke.sh n form...
n=${1:-128}; shift
mkdir -p gpurun_out
for f in "$@"; do
  EFB_ELASTIC_KERNEL=$f TUNE_REPLAY=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_elastic' -s 3 -c 1 \
    -f -o gpurun_out/ke_$f python scripts/tune_ke.py HEXA8 $n 4 > gpurun_out/ke_$f.log 2>&1
  tail -1 gpurun_out/ke_$f.log
done
