#!/bin/bash
# Scratch: A/B build variants (build_variants/*.so, selected with SPADE_LIB) over set sizes.
# usage: tools/variant_sweep.sh "<variant names>" "<set sizes>"   (default lib = "default")
cd "$(dirname "$0")/.."
for v in $1; do
  for n in $2; do
    if [ "$v" = default ]; then unset SPADE_LIB; else export SPADE_LIB=$PWD/build_variants/libspade_$v.so; fi
    echo "== variant $v N=$n"
    PYTHONPATH=$PWD python tools/quick_perf.py --N $n --S 1000 --reps 4 2>&1 | grep -E "rep[1-3]|host-output"
  done
done
