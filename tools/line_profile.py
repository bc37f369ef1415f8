#!/usr/bin/env python
"""Executed SASS instructions and stall samples per CUDA source line for one kernel:
joins `nvdisasm -g` (line info) with the ncu source page (per-instruction counters) by instruction order."""
import collections
import csv
import re
import subprocess
import sys

rep, disasm, kernel = sys.argv[1], sys.argv[2], sys.argv[3]
per = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iW = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
prof = [(r[iS].strip(), int(r[iE]), int(r[iW] or 0)) for r in rows[2:] if len(r) > iW and r[iE].isdigit()]

# walk the disassembly of that kernel: //## File "...", line N  then instructions  /*addr*/ OP ...;
lines = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith("\t.section\t.text." + kernel))
cur, seq = None, []
for l in lines[start + 1:]:
    if l.startswith("\t.section") or l.startswith("//-----"):
        if seq:
            break
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2).strip()))
assert len(seq) == len(prof), (len(seq), len(prof))
by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for (loc, txt), (src, n, w) in zip(seq, prof):
    by_line[loc][0] += n
    by_line[loc][1] += w
    by_line[loc][2][src.split()[0] if not src.startswith("@") else src.split()[1]] += n
tot = sum(v[0] for v in by_line.values())
tw = sum(v[1] for v in by_line.values())
src_cache = {}
for loc, (n, w, ops) in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:45]:
    text = ""
    if loc:
        import glob
        for f in glob.glob("/root/repo/wobble_jax_b200/csrc/" + loc[0]):
            src_cache.setdefault(f, open(f).read().splitlines())
            text = src_cache[f][loc[1] - 1].strip()[:70]
    top = ",".join(f"{k.split('.')[0]}:{v / per:.0f}" for k, v in ops.most_common(4))
    print(f"{str(loc):28s} {n / per:7.1f} {100 * n / tot:5.1f}% stall {100 * w / max(tw, 1):5.1f}%  [{top}]  {text}")
