#!/bin/bash
# dev: per-kernel timing of the fused PCG forms on small shards.  usage: gpu_pf_timing.sh <tag> <N> <pf-n> <degree>
tag=$1; N=$2; pfn=$3; deg=$4
mkdir -p gpurun_out
EFB_PCG_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29553 \
  bench.py --gpus $N --cells 32 --steps 3 --warmup 3 --no-solve --no-transient --no-parity --no-cpu --pf-config 3 --pf-n $pfn --pf-degree $deg \
  > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
grep "rank 0" gpurun_out/${tag}.err | tail -4 | cut -c1-250
python -c "
import json; d=json.load(open('gpurun_out/${tag}.json'))
v=d['extras']['phase_field_config3']; print(v.get('s_per_iter'), v.get('pcg_iters_elastic'), v.get('pcg_precond_degree'), v.get('error'))"
