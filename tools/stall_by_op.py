"""Per-opcode breakdown of a stall reason from `ncu --page source --csv --print-source sass` output.
usage: python tools/stall_by_op.py src.csv stall_short_sb [topN]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
i = 0
while rows[i][0] != "Address":
    i += 1
hdr = rows[i]
col = hdr.index(sys.argv[2])
isrc = hdr.index("Source")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 12
ops = collections.Counter()
lines = []
prev = ""
for r in rows[i + 1:]:
    if len(r) <= col or not r[col].isdigit():
        continue
    v = int(r[col])
    src = r[isrc].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else src).split(".")[0]
    ops[op] += v
    lines.append((v, src[:70], prev[:50]))
    prev = src
tot = sum(ops.values())
print(sys.argv[2], "total", tot)
for k, v in ops.most_common(top):
    print(f"  {k:10s} {v:7d} {100 * v / max(tot, 1):5.1f}%")
for v, s, p in sorted(lines, reverse=True)[:top]:
    print(f"  {v:6d}  {s}   <- after: {p}")
