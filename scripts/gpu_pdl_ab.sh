#!/bin/bash
# dev: programmatic dependent launch on/off, phase-field config 3 on N GPUs with small shards.  usage: gpu_pdl_ab.sh N pf-n
N=$1; pfn=$2
for pdl in 1 0 1 0; do
  EFB_PCG_PDL=$pdl timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 \
    bench.py --gpus $N --cells 32 --steps 3 --warmup 3 --no-solve --no-transient --no-parity --no-cpu --pf-config 3 --pf-n $pfn 2>/dev/null \
    | python -c "import json,sys; d=json.loads(sys.stdin.read()); v=d['extras']['phase_field_config3']; print('pdl $pdl', v.get('s_per_iter'), v.get('pcg_iters_elastic'), v.get('pcg_converged'), v.get('error'))"
done
