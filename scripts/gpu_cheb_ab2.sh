#!/bin/bash
# dev: per-kernel durations (EFB_PCG_TIMING) of library variants on the phase-field configs.  usage: gpu_cheb_ab2.sh variant...
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="build/variants/$v.so"
  for cfg in "3 1000" "4 119"; do
    set -- $cfg
    echo -n "$v cfg$1: "
    EFB_PCG_TIMING=1 EFB_SPMV_LANES=${SPMV_LANES:-} EFB_CHEB_LANES=${CHEB_LANES:-0} EASYFEA_B200_LIB=$lib timeout 300 python scripts/cheb_probe.py --cfg $1 --n $2 --degrees 4 2>&1 | grep "efb_pcg_iterate_cheb" | tail -30 | awk '{for(i=1;i<=NF;i++){if($i=="spmv"){s+=$(i+1)} if($i=="cheb"){c+=$(i+1)+$(i+2)+$(i+3)} } n++} END{printf "spmv %.1f us, cheb steps avg %.1f us (n=%d)\n", s/n, c/n/3, n}'
  done
done
