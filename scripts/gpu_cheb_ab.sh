#!/bin/bash
# dev: A/B of library variants (build/variants/*.so) on the phase-field configs, one GPU.  usage: gpu_cheb_ab.sh variant...
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="build/variants/$v.so"
  for cfg in "3 1000" "4 119"; do
    set -- $cfg
    echo -n "$v cfg$1: "
    EASYFEA_B200_LIB=$lib timeout 300 python scripts/cheb_probe.py --cfg $1 --n $2 --degrees 4 2>&1 | grep -v Warning | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print([i['s'] for i in d['iters']], [i['elastic'] for i in d['iters']])"
  done
done
