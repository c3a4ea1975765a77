"""Attribute the executed instructions / stall samples of one kernel in an ncu report to CUDA source lines.
usage: python tools/ncu_lines.py <report.ncu-rep> <lib.so> <mangled kernel name> [top N]
Joins `ncu --page source --csv` (SASS, per address) with `nvdisasm -g` line info of the same cubin."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
start = next(i for i, l in enumerate(dis) if l.strip().startswith(".section") and ".text." + kern in l)
line_of = {}   # offset -> (file, line, inlined-at chain text)
cur = ("?", 0)
for l in dis[start + 1:]:
    if l.strip().startswith(".section"):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        line_of[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
hdr = rows[h]
iA, iI, iS, iT = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
base = None
per_line = collections.defaultdict(lambda: [0, 0, 0])
per_op = collections.defaultdict(lambda: [0, 0])
tot_i = tot_s = tot_t = 0
for r in rows[h + 1:]:
    if len(r) < len(hdr):
        continue
    a = int(r[iA], 16)
    if base is None:
        base = a
    off = a - base
    n, s, t = int(r[iI]), int(r[iS]), int(r[iT])
    (f, ln), txt = line_of.get(off, (("?", 0), "?"))
    per_line[(f, ln)][0] += n
    per_line[(f, ln)][1] += s
    per_line[(f, ln)][2] += t
    op = txt.split()[0] if not txt.startswith("@") else txt.split()[1]
    per_op[op.split(".")[0]][0] += n
    per_op[op.split(".")[0]][1] += s
    tot_i += n
    tot_s += s
    tot_t += t
print(f"total warp instructions {tot_i}, thread instructions {tot_t} (avg {tot_t / tot_i:.1f} lanes), samples {tot_s}")
print("\n-- by source line (inst %, sample %, lanes)")
src_cache = {}
def src(f, ln):
    for d in ("branson_b200/csrc", "."):
        p = os.path.join(d, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            L = src_cache[p]
            return L[ln - 1].strip()[:90] if 0 < ln <= len(L) else ""
    return ""
for (f, ln), (n, s, t) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{100 * n / tot_i:5.1f}% {100 * s / max(tot_s, 1):5.1f}% {t / max(n, 1):5.1f}  {f}:{ln}  {src(f, ln)}")
print("\n-- by opcode (inst %, sample %)")
for op, (n, s) in sorted(per_op.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{100 * n / tot_i:5.1f}% {100 * s / max(tot_s, 1):5.1f}%  {op}")
