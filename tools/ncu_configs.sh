#!/bin/bash
# ncu --set full of one steady-state transport launch per deck: tools/ncu_configs.sh <tag> deck:skip:scale ...
#   e.g. tools/ncu_configs.sh r01 big_cube_200:1:0.2 hot_zone:3:1 marshak_wave:5:1
#   ALGO=event KREGEX=k_transport_pool tools/ncu_configs.sh r02_pool big_cube_200:1:0.2   (the event-queue kernel)
tag=$1; shift
ALGO=${ALGO:-history}
KREGEX=${KREGEX:-k_transport_history}
for spec in "$@"; do
  IFS=: read deck skip scale <<< "$spec"
  ncu --set full --import-source on --clock-control none -k regex:$KREGEX -s $skip -c 1 -f \
      -o gpurun_out/${deck}_${tag} python tools/bench_configs.py --only $deck --scale-photons $scale --algorithm $ALGO \
      > gpurun_out/ncu_${deck}_${tag}.log 2>&1
  python tools/ncu_summary.py gpurun_out/${deck}_${tag}.ncu-rep gpurun_out/${deck}_${tag}_ncu.md > /dev/null 2>&1
  cat gpurun_out/${deck}_${tag}_ncu.md
  # per-line attribution while the report is still here, then drop it unless KEEP_REP=1 (gpurun merges <= 64 MiB back)
  if [ -n "$LINES_KERNEL" ]; then
    python tools/ncu_lines.py gpurun_out/${deck}_${tag}.ncu-rep branson_b200/libbranson_gpu.so "$LINES_KERNEL" 70 \
        > gpurun_out/${deck}_${tag}_lines.txt 2>&1
  fi
  [ "$KEEP_REP" = 1 ] || rm -f gpurun_out/${deck}_${tag}.ncu-rep
done
