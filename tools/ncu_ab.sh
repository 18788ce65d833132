#!/bin/bash
# Scratch: ncu --set full of k_de_fused for layout variants (env-selected), each after a plain run.
# usage: tools/ncu_ab.sh   -> gpurun_out/prof_r2_<tag>.ncu-rep
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
run() {  # tag, env...
  tag=$1; shift
  env "$@" python tools/quick_perf.py --N 500 --S 1000 --reps 2 > gpurun_out/plain_$tag.log 2>&1 &&
  env "$@" ncu --set full --clock-control none --import-source on -k regex:k_de_fused -s 1 -c 1 \
      -o gpurun_out/prof_r2_$tag python tools/quick_perf.py --N 500 --S 1000 --reps 2 > gpurun_out/ncu_$tag.log 2>&1
  tail -2 gpurun_out/ncu_$tag.log
}
run base SPADE_TWO_CHOICE=0
run gp512 SPADE_TWO_CHOICE=0 SPADE_GENES_PER_PART=512
run gp512two SPADE_TWO_CHOICE=1 SPADE_GENES_PER_PART=512
