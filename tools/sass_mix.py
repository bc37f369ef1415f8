#!/usr/bin/env python
"""Executed-instruction mix per SASS opcode from an ncu report (source page)."""
import collections
import csv
import re
import subprocess
import sys

rep = sys.argv[1]
per = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iT, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
ops, samples = collections.Counter(), collections.Counter()
tot = 0
for r in rows[2:]:
    if len(r) <= iT or r[0].startswith("Kernel"):
        continue
    try:
        n = int(r[iE])
    except ValueError:
        continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[iS])
    op = m.group(2) if m else "?"
    ops[op] += n
    samples[op] += int(r[iW] or 0)
    tot += n
print("total warp instructions", tot, " per unit", tot / per)
ts = sum(samples.values())
for op, n in ops.most_common(28):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  per-unit {n / per:7.1f}   stall-samples {100 * samples[op] / max(ts, 1):5.1f}%")
