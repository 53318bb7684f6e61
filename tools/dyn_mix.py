#!/usr/bin/env python3
"""Dynamic SASS instruction mix per warp per walk step from an ncu report (source page)."""
import csv, subprocess, sys
from collections import Counter
rep, kern, steps = sys.argv[1], sys.argv[2], int(sys.argv[3])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1] if rows else "no rows")
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10 and r[idx['Instructions Executed']].isdigit()]
inst = []; cur = []; prev = -1
for r in data:
    a = int(r[idx['Address']], 16)
    if a < prev: inst.append(cur); cur = []
    cur.append(r); prev = a
inst.append(cur)
d = inst[0]
w = int(d[0][idx['Instructions Executed']])
c = Counter(); tot = 0
for r in d:
    n = int(r[idx['Instructions Executed']]); toks = r[idx['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    c[op.split('.')[0].rstrip(';')] += n; tot += n
print("warps", w, "instr per warp-step %.1f" % (tot / (w * steps)))
print("  ".join("%s %.1f" % (op, n / (w * steps)) for op, n in c.most_common(22)))
if len(sys.argv) > 4:
    for r in d:
        n = int(r[idx['Instructions Executed']])
        if n >= w * steps * 0.5: print("%5.2f %s" % (n / (w * steps), r[idx['Source']].strip()[:100]))
