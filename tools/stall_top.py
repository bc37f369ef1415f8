#!/usr/bin/env python
"""Top SASS instructions by stall samples (all reasons, or one given reason column) from an ncu report."""
import csv
import subprocess
import sys

rep = sys.argv[1]
col = sys.argv[2] if len(sys.argv) > 2 else "# Samples"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iC, iE = hdr.index("Source"), hdr.index(col), hdr.index("Instructions Executed")
data = []
for k, r in enumerate(rows[2:]):
    if len(r) > iC and r[iC].isdigit():
        data.append((int(r[iC]), k, r[iS].strip(), int(r[iE])))
tot = sum(d[0] for d in data)
print(col, "total", tot)
for c, k, s, e in sorted(data, reverse=True)[:n]:
    prev = data[k - 1][2][:50] if k > 0 else ""
    print(f"{100 * c / tot:5.1f}%  #{k:5d} x{e:8d}  {s[:70]:70s} | prev: {prev}")
