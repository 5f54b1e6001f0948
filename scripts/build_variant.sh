#!/bin/bash
# dev tool: build a library variant with extra nvcc flags into build/variants/<name>.so   (usage: build_variant.sh name [flags...])
set -e
cd "$(dirname "$0")/.."
name=$1; shift
tmp=$(mktemp -d)
for f in api elem_kernels pf_kernels csr_kernels pcg_kernels post_kernels fused_kernels fused_mma; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v "$@" -c easyfea_b200/csrc/$f.cu -o $tmp/$f.o 2> $tmp/$f.log &
done
wait
mkdir -p build/variants
nvcc -shared -o build/variants/$name.so $tmp/*.o -cudart static 2>/dev/null
grep -A2 "k_elasticILi3ELi8ELi0" $tmp/elem_kernels.log | grep -E "Used|spill" || true
rm -rf $tmp
