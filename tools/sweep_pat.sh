#!/bin/bash
# usage: sweep_pat.sh corpus gib pattern variant...
cfg=$1; gib=$2; pat=$3; shift 3
for v in "$@"; do
  echo -n "$v $pat: "
  UGX_LIB=ugrep_b200/build/$v.so python tools/prof_one.py --config $cfg --gib $gib --reps 5 --pattern tests/golden/patterns/$pat.ugxp --mode lines 2>&1 | cut -c1-80
done
