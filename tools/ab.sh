#!/bin/bash
# A/B timing of library variants on the bench workload: tools/ab.sh <photons> main nore0 nolog ...
# ("main" = the in-tree build; others = build/variants/<name>, see `make -C branson_b200/csrc variant`)
photons=$1; shift
for v in "$@"; do
  if [ "$v" = main ]; then unset BRANSON_LIB_DIR; else export BRANSON_LIB_DIR=$PWD/build/variants/$v; fi
  python bench.py --no-cpu-baseline --photons $photons --steps 4 --warmup 3 > gpurun_out/ab_$v.log 2>&1
  python - "$v" gpurun_out/ab_$v.log <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(f"{sys.argv[1]:>10}: value {j['value']/1e6:8.2f} M/s  kernel {j['roofline']['kernel_ms_per_launch']:8.3f} ms  e2e {j['e2e']['value']/1e6:8.2f} M/s")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
