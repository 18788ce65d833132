#!/bin/bash
# Scratch: A/B layout variants through the environment (SPADE_TWO_CHOICE, SPADE_GENES_PER_PART).
cd "$(dirname "$0")/.."
for tc in 0 1; do for gp in 1024 768 512; do for n in $1; do
  echo "== two_choice=$tc Gp=$gp N=$n"
  SPADE_TWO_CHOICE=$tc SPADE_GENES_PER_PART=$gp PYTHONPATH=$PWD python tools/quick_perf.py --N $n --S 1000 --reps 4 2>&1 | grep -E "rep[1-3]|stage"
done; done; done
