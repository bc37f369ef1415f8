#!/usr/bin/env python
"""Static opcode mix of an address range of a kernel's SASS (cuobjdump -sass output).
usage: sass_static.py file.sass [lo_index hi_index]"""
import collections, re, sys
ops = []
for l in open(sys.argv[1]):
    m = re.match(r'\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);', l)
    if m:
        ops.append((int(m.group(1), 16), m.group(2).strip()))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(ops)
c = collections.Counter()
for a, t in ops[lo:hi]:
    t = re.sub(r'^@!?U?P\d+\s+', '', t)
    c[t.split()[0].split('.')[0]] += 1
print("instructions", hi - lo)
fp64 = sum(v for k, v in c.items() if k in ("DFMA", "DADD", "DMUL"))
print("fp64", fp64)
print(" ".join(f"{k}:{v}" for k, v in c.most_common(40)))
