#!/bin/bash
# dev: selected GPU tests + a short bench.  usage: gpu_quick.sh "<pytest args>" "<bench args>"
mkdir -p gpurun_out
python -m pytest $1 -x -q 2>&1 | tail -5
if [ -n "$2" ]; then python bench.py $2 > gpurun_out/quick_bench.json 2> gpurun_out/quick_bench.err; tail -3 gpurun_out/quick_bench.err; fi
