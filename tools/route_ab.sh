#!/bin/bash
# usage: route_ab.sh corpus gib pattern...   default route vs the streaming DFA route
cfg=$1; gib=$2; shift 2
for p in "$@"; do
  echo -n "default $p: "; python tools/prof_one.py --config $cfg --gib $gib --reps 5 --pattern tests/golden/patterns/$p.ugxp --mode lines 2>&1 | cut -c1-70
  echo -n "stream  $p: "; python tools/prof_one.py --config $cfg --gib $gib --reps 5 --pattern tests/golden/patterns/$p.ugxp --mode lines --opt stream_dfa=1 2>&1 | cut -c1-70
done
