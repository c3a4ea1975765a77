#!/bin/bash
# A/B of library variants on the per-deck sweep, same box, interleaved twice: tools/ab_configs.sh "<decks>" <scale> main v7 cut ...
decks=$1; scale=$2; shift 2
for rep in 1 2; do
for v in "$@"; do
  if [ "$v" = main ]; then unset BRANSON_LIB_DIR; else export BRANSON_LIB_DIR=$PWD/build/variants/$v; fi
  echo "#### $v (rep $rep)"
  python tools/bench_configs.py --only $decks --scale-photons $scale 2>&1 | grep -v "^  cyc" | awk '/^==/ {name=$2} /^ +[0-9]/ {last[name]=$4; ms[name]=$3} END {for (n in last) printf "   %-28s last-cycle %8.1f Mhist/s (%s ms)\n", n, last[n], ms[n]}'
done
done
