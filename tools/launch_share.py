#!/usr/bin/env python
"""Per-kernel count / mean / total / share from an ncu launch list (`--metrics gpu__time_duration.sum --csv`).
    python tools/launch_share.py profiles/launches_r02.csv [skip_first_n]"""
import csv
import re
import sys
from collections import OrderedDict

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
rows = list(csv.DictReader(lines))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = [r for r in rows if r.get("Metric Name") == "gpu__time_duration.sum"][skip:]
agg = OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("<unnamed>::", "").replace("void ", "")
    v = float(r["Metric Value"].replace(",", ""))
    if r["Metric Unit"] in ("ns", "nsecond"):
        v /= 1000.0
    elif r["Metric Unit"] in ("ms", "msecond"):
        v *= 1000.0
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':60s} {'n':>6s} {'mean_us':>9s} {'total_us':>10s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} {n:6d} {t / n:9.2f} {t:10.1f} {t / tot:6.3f}")
