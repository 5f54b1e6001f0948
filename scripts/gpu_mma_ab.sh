#!/bin/bash
# dev tool: A/B of k_assemble_hexa8_mma library variants (build/variants/*.so made by build_variant.sh) on one box
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in default "$@"; do
  if [ "$v" = default ]; then unset EASYFEA_B200_LIB; else export EASYFEA_B200_LIB=$PWD/build/variants/$v.so; fi
  echo "== $v"
  timeout 90 python scripts/fused_probe.py --n 128 --S "" --reps 10 2>&1 | tail -1
done | tee gpurun_out/mma_ab.log
