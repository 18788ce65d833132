#!/bin/bash
# Scratch: A/B of one environment switch of the engine over set sizes.
# usage: tools/env_ab.sh VAR "<values>" "<set sizes>" [quick_perf args]
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
var=$1; vals=$2; sizes=$3; shift 3
for n in $sizes; do
  for v in $vals; do
    echo "== N=$n $var=$v"
    env $var=$v python tools/quick_perf.py --N $n --S 1000 --reps 5 "$@" 2>&1 | grep -E "rep[2-4]"
  done
done
