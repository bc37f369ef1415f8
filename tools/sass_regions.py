#!/usr/bin/env python
"""Group the SASS rows of an ncu source page into runs of equal execution count (= code regions):
instructions, executed warp instructions, stall samples per region, plus the top stall reasons.
usage: sass_regions.py report.ncu-rep [min_share_percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") or h.startswith("Stall")]
data = []
for r in rows[2:]:
    try:
        data.append((r[iS].strip(), int(r[iE]), int(r[iW] or 0), r))
    except (ValueError, IndexError):
        pass
tot_e = sum(d[1] for d in data)
tot_s = sum(d[2] for d in data)
print(f"{len(data)} SASS instructions, {tot_e} executed, {tot_s} samples")
# regions: consecutive rows whose execution count is within 2 % of the region's first row
regions = []
for k, d in enumerate(data):
    if regions and abs(d[1] - regions[-1]["cnt"]) <= 0.02 * max(regions[-1]["cnt"], 1):
        g = regions[-1]
    else:
        g = {"start": k, "cnt": d[1], "n": 0, "e": 0, "s": 0, "ops": {}}
        regions.append(g)
    g["n"] += 1
    g["e"] += d[1]
    g["s"] += d[2]
    op = d[0].split()[0] if not d[0].startswith("@") else d[0].split()[1]
    op = op.split(".")[0]
    g["ops"][op] = g["ops"].get(op, 0) + 1
for g in regions:
    if 100 * g["e"] / tot_e < min_share and 100 * g["s"] / max(tot_s, 1) < min_share:
        continue
    ops = ", ".join(f"{o}:{c}" for o, c in sorted(g["ops"].items(), key=lambda kv: -kv[1])[:8])
    print(f"#{g['start']:5d} n={g['n']:4d} x{g['cnt']:9d}  inst {100 * g['e'] / tot_e:5.1f}%  samples {100 * g['s'] / max(tot_s, 1):5.1f}%  {ops}")
