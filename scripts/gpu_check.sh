#!/bin/bash
# Run on the B200 box (via gpurun): GPU parity tests, the bench line, an ncu launch list of a short bench run and one
# `--set full` capture of the two dominant kernels.  Everything lands in gpurun_out/.
#   usage: scripts/gpu_check.sh [tag] [bench-n] [what...]      what ⊂ {tests bench launches full}
tag=${1:-r1}
n=${2:-200}
shift 2 || true
what=${*:-tests bench launches full}
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt 2>&1
for w in $what; do
  case $w in
    tests)
      timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
      echo "pytest exit $?" >> gpurun_out/${tag}_pytest.log
      tail -5 gpurun_out/${tag}_pytest.log ;;
    smoke)
      timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
      tail -2 gpurun_out/${tag}_smoke.log ;;
    bench)
      timeout 1500 python bench.py --n $n > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
      tail -c 3000 gpurun_out/${tag}_bench.json; tail -5 gpurun_out/${tag}_bench.err ;;
    launches)
      timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_LAUNCHES:-400} --csv \
        --log-file gpurun_out/${tag}_launches.csv python bench.py --n 128 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 --no-pf --no-solve --no-transient \
        > gpurun_out/${tag}_launches.log 2>&1
      tail -2 gpurun_out/${tag}_launches.log ;;
    full)
      timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-k_elastic|k_replay|k_assemble}" -s ${NCU_SKIP:-6} -c ${NCU_COUNT:-2} \
        -f -o gpurun_out/${tag}_full python bench.py --n 128 --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-pf --no-solve --no-transient \
        > gpurun_out/${tag}_full.log 2>&1
      tail -2 gpurun_out/${tag}_full.log ;;
  esac
done
