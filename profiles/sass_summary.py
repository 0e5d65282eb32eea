"""Summarise an `ncu --page source --csv --print-source sass` dump: executed warp-instructions by
opcode and the warp-stall sample breakdown, per kernel.  Usage: python profiles/sass_summary.py file.csv [nwarps]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
nwarps = float(sys.argv[2]) if len(sys.argv) > 2 else None
kern, hdr = None, None
acc = {}
for r in rows:
    if r and r[0] == "Kernel Name":
        kern = r[1]
        acc.setdefault(kern, dict(ops=collections.Counter(), stalls=collections.Counter(), n=0, sass=0))
        hdr = None
        continue
    if r and r[0] == "Address":
        hdr = {h: i for i, h in enumerate(r)}
        stall_cols = [h for h in r if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    a = acc[kern]
    src = r[hdr["Source"]].strip()
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = (m.group(2) if m else src[:10]).split(".")[0]
    n = int(r[hdr["Instructions Executed"]] or 0)
    a["ops"][op] += n
    a["n"] += n
    a["sass"] += 1
    for c in stall_cols:
        a["stalls"][c] += int(r[hdr[c]] or 0)
seen = set()
for k, a in acc.items():
    key = (k, a["n"])
    if key in seen:
        continue
    seen.add(key)
    print(f"== {k}\n   SASS instructions {a['sass']}, executed warp-instructions {a['n']}" +
          (f", per warp {a['n'] / nwarps:.0f}" if nwarps else ""))
    for op, v in a["ops"].most_common(24):
        print(f"   {op:12s} {v:12d} {100 * v / max(a['n'], 1):5.1f}%")
    ts = sum(a["stalls"].values())
    print("   stall samples %:", {c: round(100 * v / max(ts, 1), 1) for c, v in a["stalls"].most_common(9)})
