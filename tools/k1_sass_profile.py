#!/usr/bin/env python
"""Per-SASS-instruction execution counts of one kernel from an `ncu --set full --import-source on` report, normalised to
executions per unit of work (K1: per warp-step = one image row of one 256-pixel strip), split at CALL/RET boundaries.
  python tools/k1_sass_profile.py <report.ncu-rep> <units> [listing-out]
"""
import collections
import csv
import io
import subprocess
import sys

rep, units = sys.argv[1], float(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
for i, r in enumerate(rows):
    if r and r[0] == "Kernel Name":
        H, start = rows[i + 1], i + 2
        break
iS, iI, iN = H.index("Source"), H.index("Instructions Executed"), H.index("# Samples")
L = []
for r in rows[start:]:
    if not r or r[0] == "Kernel Name":
        break
    L.append((int(r[iI] or 0) / units, int(r[iN] or 0), r[iS].strip()))
tot = sum(x[0] for x in L)
print(f"{len(L)} SASS instructions, {tot:.1f} warp-instructions per unit, issue samples {sum(x[1] for x in L)}")
# functions: split after EXIT following the main body / at RET
bounds = [i for i, x in enumerate(L) if x[2].startswith("RET") or (x[2].startswith("EXIT") and x[0] > 0)]
prev = 0
for b in bounds:
    print(f"  instructions {prev:5d}..{b:5d}: {sum(x[0] for x in L[prev:b + 1]):7.1f} per unit")
    prev = b + 1
ops = collections.Counter()
for c, n, t in L:
    p = t.split()
    op = p[1] if p[0].startswith("@") else p[0]
    ops[op.split(".")[0]] += c
print("  " + "  ".join(f"{k} {v:.1f}" for k, v in ops.most_common(24)))
if len(sys.argv) > 3:
    with open(sys.argv[3], "w") as f:
        acc = 0.0
        for c, n, t in L:
            acc += c
            f.write(f"{acc:8.2f} {c:7.3f} {n:6d}  {t}\n")
