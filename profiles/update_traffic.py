"""Record the measured DRAM traffic of a sweep kernel in profiles/traffic.json, keyed by workload, precision and
kernel and stamped with the hash of the kernel sources it was captured on (bench.py quotes it as roofline.traffic
only while that hash is current).

    python profiles/update_traffic.py <report.ncu-rep> <workload:precision:kernel> <sweeps per captured launch> [scale]

`scale` multiplies the per-launch bytes (a capture taken on a fraction of the bench batch: 4 for 262144 of 1048576
cases).  Takes the LAST captured launch of the report."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import kernel_source_hash  # noqa: E402

rep, key, sweeps = sys.argv[1], sys.argv[2], float(sys.argv[3])
scale = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units, last = rows[0], rows[1], rows[-1]


def val(name):
    i = hdr.index(name)
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
    return float(last[i]) * mult


total = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) * scale / sweeps
path = os.path.join(ROOT, "profiles", "traffic.json")
tj = json.load(open(path)) if os.path.exists(path) else {}
tj.setdefault("entries", {})[key] = {"dram_bytes_per_sweep": total, "source_hash": kernel_source_hash(),
                                     "from": os.path.relpath(rep, ROOT), "kernel": last[hdr.index("Kernel Name")][:60]}
json.dump(tj, open(path, "w"), indent=1)
print(key, total, "bytes per sweep")
