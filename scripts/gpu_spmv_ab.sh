#!/bin/bash
# dev: A/B builds of the SpMV (prefetch depth)
for cfg in "HEXA8 100 3" "TRI3 1000 2" "TRI3 1000 1" "TETRA4 60 3" "HEXA8 60 1"; do
  set -- $cfg
  for lib in "" build/variants/pf3.so build/variants/pf4.so; do
    echo -n "lib=${lib:-default} "; EASYFEA_B200_LIB=$lib PROBE_ELEM=$1 PROBE_N=$2 PROBE_DOF=$3 timeout 120 python scripts/spmv_probe.py 2>&1 | tail -1
  done
done
timeout 300 python -m pytest tests/test_gpu_phasefield_solver.py tests/test_gpu_staggered.py -m gpu -x -q 2>&1 | tail -3
PROBE_ITERS=200 timeout 120 python scripts/pcg_probe.py; PROBE_ELEM=TRI3 PROBE_N=1000 PROBE_ITERS=500 timeout 120 python scripts/pcg_probe.py
