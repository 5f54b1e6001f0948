#!/bin/bash
# dev: phase-field legs only on N GPUs for several polynomial degrees.  usage: gpu_pf8.sh <tag> <N> <degree>...
tag=$1; N=$2; shift 2
mkdir -p gpurun_out
for deg in "$@"; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 \
    bench.py --gpus $N --cells 48 --steps 3 --warmup 3 --no-solve --no-transient --no-parity --no-cpu --pf-degree $deg \
    > gpurun_out/${tag}_deg${deg}.json 2> gpurun_out/${tag}_deg${deg}.err
  echo "degree $deg exit $?"; python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_deg${deg}.json"))
for k,v in d["extras"].items():
    if k.startswith("phase"): print(k, v.get("s_per_iter"), v.get("pcg_iters_elastic"), v.get("pcg_iters_damage"), v.get("pcg_converged"), v.get("pcg_precond_degree"), v.get("error"))
PY
done
