"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares, then the sequence.
usage: python tools/launch_list.py gpurun_out/launches.csv profiles/launches_rNN.csv "<command that was profiled>" """
import collections
import csv
import re
import sys

src, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
h = rows[0]
iK, iM, iV, iU = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
seq = []
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    v = float(r[iV].replace(",", ""))
    v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iU], 1)
    name = re.sub(r"\(.*", "", r[iK]).replace("void ", "")
    seq.append((name, v))
tot = sum(v for _, v in seq)
agg = collections.OrderedDict()
for n, v in seq:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
with open(out, "w") as f:
    f.write(f"# ncu launch list: `{cmd}`\n# (times are cold-cache and serialised: compare SHARES, not absolutes)\n")
    f.write("kernel,launches,total_ns,share\n")
    for n, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{n},{c},{v:.0f},{v / tot:.4f}\n")
    f.write("\n# per-launch sequence (kernel,ns)\n")
    for n, v in seq:
        f.write(f"{n},{v:.0f}\n")
print(open(out).read()[:1500])
