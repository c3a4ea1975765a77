#!/bin/bash
# ncu counters (not timings) of the steady-state transport launch for library variants: tools/ncu_ab.sh <photons> main base ...
photons=$1; shift
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__sass_inst_executed_op_local_ld.sum,smsp__sass_inst_executed_op_local_st.sum,smsp__pcsamp_warps_issue_stalled_wait,smsp__pcsamp_warps_issue_stalled_no_instructions,smsp__pcsamp_warps_issue_stalled_long_scoreboard,smsp__pcsamp_warps_issue_stalled_short_scoreboard,smsp__pcsamp_warps_issue_stalled_math_pipe_throttle,smsp__pcsamp_warps_issue_stalled_not_selected,smsp__pcsamp_warps_issue_stalled_selected,smsp__pcsamp_warps_issue_stalled_barrier,smsp__pcsamp_warps_issue_stalled_branch_resolving,smsp__pcsamp_warps_issue_stalled_dispatch_stall,smsp__pcsamp_warps_issue_stalled_lg_throttle,smsp__pcsamp_warps_issue_stalled_imc_miss
for v in "$@"; do
  if [ "$v" = main ]; then unset BRANSON_LIB_DIR; else export BRANSON_LIB_DIR=$PWD/build/variants/$v; fi
  ncu --metrics $M --clock-control none -k regex:k_transport_history -s 3 -c 1 --csv --log-file gpurun_out/ncuab_$v.csv \
      python bench.py --no-cpu-baseline --photons $photons --steps 1 --warmup 3 > gpurun_out/ncuab_$v.log 2>&1
  echo "== $v"; python - gpurun_out/ncuab_$v.csv <<'PY'
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; i_n = h.index("Metric Name"); i_v = h.index("Metric Value")
d = {r[i_n]: r[i_v] for r in rows[1:]}
st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v.replace(",", "")) for k, v in d.items() if "pcsamp" in k and v != "n/a"}
tot = sum(st.values()) or 1
for k, v in d.items():
    if "pcsamp" not in k: print(f"  {k}: {v}")
print("  stalls:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])))
PY
done
