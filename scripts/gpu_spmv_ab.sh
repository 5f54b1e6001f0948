#!/bin/bash
# dev: A/B the lanes-per-node choice of the SpMV
for cfg in "TETRA4 80 3" "TETRA4 80 1" "TRI3 1000 2" "TRI3 1000 1" "HEXA8 100 3" "HEXA8 100 1" "QUAD9 500 2" "HEXA27 30 3"; do
  set -- $cfg
  for l in 4 8 16 32; do
    echo -n "lanes=$l "; EFB_SPMV_LANES=$l PROBE_ELEM=$1 PROBE_N=$2 PROBE_DOF=$3 timeout 120 python scripts/spmv_probe.py 2>&1 | tail -1
  done
done
