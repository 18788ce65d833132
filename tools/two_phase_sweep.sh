#!/bin/bash
# Scratch: fused finalize (SPADE_TWO_PHASE=0) against raw words + k_expand (1) over set sizes.
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
for n in $1; do
  for tp in 0 1; do
    echo "== N=$n two_phase=$tp"
    SPADE_TWO_PHASE=$tp python tools/quick_perf.py --N $n --S ${2:-1000} --reps 4 2>&1 | grep -E "rep[1-3]"
  done
done
