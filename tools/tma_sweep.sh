#!/bin/bash
# Scratch: k_de_tma (segments staged through a shared-memory ring: SPADE_TMA=1 cp.async.bulk, 2 cp.async)
# against k_de_fused, single and two-choice accumulators.
# usage: tools/tma_sweep.sh "<variant names>" "<set sizes>" "<tma:two_choice modes>"   (variant "default" = the product library)
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
for v in $1; do
  for n in $2; do
    for mode in $3; do
      tma=${mode%:*}; two=${mode#*:}
      if [ "$v" = default ]; then unset SPADE_LIB; else export SPADE_LIB=$PWD/build_variants/libspade_$v.so; fi
      echo "== variant $v N=$n tma=$tma two_choice=$two"
      SPADE_TMA=$tma SPADE_TWO_CHOICE=$two python tools/quick_perf.py --N $n --S 1000 --reps 4 2>&1 | grep -E "rep[1-3]"
    done
  done
done
