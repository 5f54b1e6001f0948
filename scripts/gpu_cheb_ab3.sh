#!/bin/bash
# dev: per-kernel durations on config 3 (TRI3 d=2 elastic + d=1 damage) for library variants x lanes.  usage: gpu_cheb_ab3.sh variant...
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="build/variants/$v.so"
  for l in 2 4; do
    echo -n "$v lanes $l: "
    EFB_PCG_TIMING=1 EFB_SPMV_LANES=$l EFB_CHEB_LANES=$l EASYFEA_B200_LIB=$lib timeout 300 python scripts/cheb_probe.py --cfg 3 --n 1000 --degrees 4 2>&1 | grep "efb_pcg_iterate_cheb" | awk '{for(i=1;i<=NF;i++){if($i=="spmv"){s=$(i+1)} if($i=="cheb"){c=($(i+1)+$(i+2)+$(i+3))/3} } if (s+0 > 60) {se+=s; ce+=c; ne++} else {sd+=s; cd+=c; nd++}} END{printf "elastic spmv %.1f cheb %.1f (n=%d) | damage spmv %.1f cheb %.1f (n=%d)\n", se/ne, ce/ne, ne, sd/nd, cd/nd, nd}'
  done
done
