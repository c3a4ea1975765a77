"""Summarise an ncu report (one kernel launch per row of --page raw) into markdown + a traffic json.
usage: python tools/ncu_summary.py gpurun_out/transport_r01.ncu-rep profiles/transport_r01_ncu.md [traffic.json]"""
import csv
import json
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__inst_executed_op_global_red.sum", "smsp__sass_inst_executed_op_local_ld.sum",
    "smsp__sass_inst_executed_op_local_st.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_op_read_hit_rate.pct", "lts__t_sector_op_red_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum", "lts__average_t_sector_hit_rate_realtime.pct",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_op_shared_st.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.per_cycle_active", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "smsp__inst_executed.avg.per_cycle_active",
]
lines = [f"# ncu --set full summary of `{rep}`", ""]
traffic = []
for r in rows[2:]:
    if len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    u = dict(zip(hdr, units))
    lines.append(f"## {d.get('Kernel Name', '?')}  (launch id {d.get('ID', '?')})")
    lines.append("")
    lines.append("| metric | value | unit |")
    lines.append("|---|---|---|")
    for k in KEYS:
        if k in d:
            lines.append(f"| {k} | {d[k]} | {u[k]} |")
    st = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v) for k, v in d.items()
          if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v not in ("", "n/a")}
    tot = sum(st.values()) or 1.0
    lines.append("")
    lines.append("warp-state samples: " + ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in
                                                    sorted(st.items(), key=lambda kv: -kv[1])[:9]))
    lines.append("")

    def gb(x, unit):
        x = float(x)
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}[unit]
    try:
        traffic.append(gb(d["dram__bytes_read.sum"], u["dram__bytes_read.sum"]) +
                       gb(d["dram__bytes_write.sum"], u["dram__bytes_write.sum"]))
    except Exception:
        pass
open(out, "w").write("\n".join(lines) + "\n")
if len(sys.argv) > 3 and traffic:
    json.dump({"source": rep, "kernel": "k_transport_history", "dram_bytes_per_launch": sum(traffic) / len(traffic),
               "launches": len(traffic)}, open(sys.argv[3], "w"))
print("\n".join(lines[:40]))
