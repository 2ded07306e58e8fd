#!/usr/bin/env python3
"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.OrderedDict()
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    d = dict(zip(hdr, r))
    if d.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = d["Kernel Name"].split("(")[0][:70]
    v = float(d["Metric Value"].replace(",", ""))
    unit = d["Metric Unit"]
    if unit in ("nsecond", "ns"):
        v /= 1000.0
    elif unit in ("msecond", "ms"):
        v *= 1000.0
    agg.setdefault(name, []).append(v)
for n, v in agg.items():
    print(f"{n:70s} n={len(v):3d} last={v[-1]:9.1f} us  min={min(v):9.1f}")
