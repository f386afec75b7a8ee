"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time and share.
    python tools/ncu_launches.py gpurun_out/launches.csv [steps_captured]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ix = {k: i for i, k in enumerate(hdr)}
agg = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*$", "", r[ix["Kernel Name"]]).replace("void ", "").replace("<unnamed>::", "").strip()
    a = agg[name]
    a[0] += 1
    a[1] += float(r[ix["Metric Value"]]) / 1e6      # ns -> ms
total = sum(v[1] for v in agg.values())
print(f"{len(rows) - 1} launches captured, {total:.3f} ms of kernel time ({total / steps:.3f} ms per step over {steps:g} steps; ncu times are cold-cache, serialised)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>9s} {'share':>7s}")
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} {n:8d} {ms:9.3f} {100 * ms / total:6.1f}%")
