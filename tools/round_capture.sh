#!/bin/bash
# One GPU call that produces the evidence of a kernel version: tools/round_capture.sh <tag> [configs]
#   gpurun_out/bench_<tag>_n1.json        the contract bench line (timed outside any profiler)
#   gpurun_out/transport_<tag>.ncu-rep    ncu --set full of one steady-state transport launch of the same command
#   gpurun_out/launches_<tag>_raw.csv     the launch list (gpu__time_duration.sum per kernel)
#   gpurun_out/configs_<tag>.{json,txt}   every BASELINE deck, per cycle (when "configs" is given)
tag=$1
python bench.py > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err
tail -c 600 gpurun_out/bench_${tag}_n1.json | head -c 300; echo
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-host-mesh-e2e --no-aos-dropin --no-parity-check"
ncu --set full --import-source on --clock-control none -k regex:k_transport_history -s 3 -c 1 -f \
    -o gpurun_out/transport_${tag} $BENCH > gpurun_out/ncu_transport_${tag}.log 2>&1
python tools/ncu_summary.py gpurun_out/transport_${tag}.ncu-rep gpurun_out/transport_${tag}_ncu.md gpurun_out/transport_traffic_${tag}.json > /dev/null 2>&1
python tools/ncu_lines.py gpurun_out/transport_${tag}.ncu-rep branson_b200/libbranson_gpu.so _ZN2bg19k_transport_historyILi0ELb0ELb1ELb0ELb1EEEvNS_15TransportParamsE 90 > gpurun_out/transport_${tag}_lines.txt 2>&1
[ "$KEEP_REP" = 1 ] || rm -f gpurun_out/transport_${tag}.ncu-rep
head -40 gpurun_out/transport_${tag}_ncu.md
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_raw.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-host-mesh-e2e --no-aos-dropin --no-parity-check > gpurun_out/launches_${tag}.log 2>&1
if [ "$2" = configs ]; then
  python tools/bench_configs.py --out gpurun_out/configs_${tag}.json > gpurun_out/configs_${tag}.txt 2>&1
  grep -E "^==|^ +[0-9]+ " gpurun_out/configs_${tag}.txt | tail -45
fi
