"""Per-source-line executed instructions / stall samples from
`ncu -i rep --page source --csv --print-source cuda,sass`.  Usage: python profiles/line_summary.py rep.ncu-rep [top]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
cur_file, cur_fn, hdr = None, None, None
agg = {}
first_fn = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        first_fn = first_fn or cur_fn
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or cur_fn != first_fn or not r[0].strip():
        continue
    try:
        n = int(r[hdr["Instructions Executed"]])
        s = int(r[hdr["# Samples"]])
    except (ValueError, KeyError):
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, 0, r[1].strip()[:90]])
    a[0] += n
    a[1] += s
tot = sum(a[0] for a in agg.values())
tots = sum(a[1] for a in agg.values())
print(first_fn)
print(f"total executed warp-instructions {tot}, samples {tots}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} {100 * a[0] / tot:5.1f}% inst {100 * a[1] / max(tots, 1):5.1f}% smpl  {a[2]}")
