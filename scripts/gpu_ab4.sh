#!/bin/bash
# dev: sweep library variants for the replay kernel + one ncu capture.  usage: gpu_ab4.sh n form variant...
n=${1:-160}; form=$2; shift 2
mkdir -p gpurun_out
for v in default "$@"; do
  if [ "$v" = default ]; then unset EASYFEA_B200_LIB; else export EASYFEA_B200_LIB=$PWD/build/variants/$v.so; fi
  EFB_REPLAY_KERNEL=$form TUNE_REPLAY=1 python scripts/tune_ke.py HEXA8 $n 10 2>&1 | tail -1 | sed "s/^/$v /" | tee -a gpurun_out/ab4.log
done
unset EASYFEA_B200_LIB
EFB_REPLAY_KERNEL=$form TUNE_REPLAY=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_replay' -s 3 -c 1 \
    -f -o gpurun_out/replay_$form python scripts/tune_ke.py HEXA8 128 4 > gpurun_out/replay_$form.log 2>&1
