#!/bin/bash
# tools/sweep_tuning.sh "<decks>" <scale> VAR=val[,VAR=val] ...   (each spec = one run of tools/bench_configs.py)
decks=$1; scale=$2; shift 2
for spec in "$@"; do
  echo "#### $spec"
  env $(echo $spec | tr ',' ' ') python tools/bench_configs.py --only $decks --scale-photons $scale 2>&1 | grep -v "^  cyc"
done
