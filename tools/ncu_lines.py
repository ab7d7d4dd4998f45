#!/usr/bin/env python
"""Attribute ncu per-SASS-instruction counts to CUDA source lines.
  python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> <cubin> [top] [mangled-substring]
(the last argument picks one template instance in the cubin, e.g. epipolar_kernelILb0ELb1EE)
Uses `ncu --page source --csv` (SASS view: executed instructions + stall samples per instruction) and
`nvdisasm -g -c` (line info) on the cubin that holds the kernel; both list the function's instructions in order.
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kname, cubin = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cname = sys.argv[5] if len(sys.argv) > 5 else kname
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
# sections: "Kernel Name",<name> / header / rows ...
sass = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name" and kname in rows[i][1] and not sass:
        H = rows[i + 1]
        iS, iI, iN = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
        j = i + 2
        while j < len(rows) and rows[j] and rows[j][0] != "Kernel Name":
            sass.append((rows[j][iS].strip(), int(rows[j][iI] or 0), int(rows[j][iN] or 0)))
            j += 1
        i = j
    else:
        i += 1
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
start = [k for k, l in enumerate(dis) if l.startswith(".text.") and cname in l][0]
lines = []
cur = None
for l in dis[start + 1:]:
    if l.startswith("//---") or l.startswith(".L_x_") and False:
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
    if l.startswith("\t.section") or (l.startswith(".text.") and cname not in l):
        break
n = min(len(lines), len(sass))
agg = collections.defaultdict(lambda: [0, 0])
for k in range(n):
    a = agg[lines[k]]
    a[0] += sass[k][1]
    a[1] += sass[k][2]
ti = sum(v[0] for v in agg.values()) or 1
ts = sum(v[1] for v in agg.values()) or 1
print(f"# {kname}: {len(sass)} SASS instructions in report, {len(lines)} in cubin; total warp-instr {ti}, samples {ts}")
src = {}
key = (lambda kv: -kv[1][1]) if (len(sys.argv) > 5 and sys.argv[5] == "stall") else (lambda kv: -kv[1][0])
for (f, ln), v in sorted(agg.items(), key=key)[:top]:
    if f not in src:
        try:
            src[f] = open(subprocess.run(["bash", "-c", f"ls /root/repo/srrg2_proslam_b200/csrc/{f} /root/repo/include/{f} 2>/dev/null | head -1"],
                                         capture_output=True, text=True).stdout.strip()).read().splitlines()
        except Exception:
            src[f] = []
    text = src[f][ln - 1].strip() if 0 < ln <= len(src[f]) else ""
    print(f"{100 * v[0] / ti:5.1f}% inst {100 * v[1] / ts:5.1f}% stall-samples  {f}:{ln:<4d} {text[:100]}")
