"""Per-kernel summary of an ncu launch list (`--metrics gpu__time_duration.sum --csv --log-file x.csv`).

    python scripts/launch_summary.py gpurun_out/launches.csv > profiles/rN_launch_summary.txt
"""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
per = OrderedDict()
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]])
    if r[ix["Metric Unit"]] in ("ns", "nsecond"):
        v /= 1e3
    per.setdefault(r[ix["Kernel Name"]], []).append(v)
tot = sum(sum(v) for v in per.values())
print(f"{'kernel':68s} {'n':>5s} {'mean us':>9s} {'min':>8s} {'max':>8s} {'share':>6s}")
for k, v in per.items():
    print(f"{k[:68]:68s} {len(v):5d} {sum(v) / len(v):9.1f} {min(v):8.1f} {max(v):8.1f} {100 * sum(v) / tot:5.1f}%")
