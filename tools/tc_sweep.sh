#!/bin/bash
# Scratch: two-choice layout with deeper member batches at reduced occupancy.
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
for cfg in "b8c3 1024" "b4c3 1024" "b8c4 768" "b8c3 768"; do
  set -- $cfg
  echo "== variant $1 two_choice=1 Gp=$2"
  SPADE_LIB=$PWD/build_variants/libspade_$1.so SPADE_TWO_CHOICE=1 SPADE_GENES_PER_PART=$2 python tools/quick_perf.py --N 500 --S 1000 --reps 4 2>&1 | grep -E "rep[1-3]"
done
echo "== variant b8c3 two_choice=0 Gp=2048"
SPADE_LIB=$PWD/build_variants/libspade_b8c3.so SPADE_TWO_CHOICE=0 SPADE_GENES_PER_PART=2048 python tools/quick_perf.py --N 500 --S 1000 --reps 4 2>&1 | grep -E "rep[1-3]"
