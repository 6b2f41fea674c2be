#!/bin/bash
# usage: tools/sweep.sh <config> <gib> name1 name2 ...   (variants built by tools/variants.py)
cfg=$1; gib=$2; shift 2
for v in "$@"; do
  echo -n "$v: "
  UGX_LIB=ugrep_b200/build/$v.so python tools/prof_one.py --config $cfg --gib $gib --reps 10 2>&1 | cut -c1-90
done
