#!/usr/bin/env python3
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: total ms, launches, share.
    python tools/launch_summary.py gpurun_out/launches.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[start:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3}.get(r[ui], 1.0)
    a = agg.setdefault(r[ki][:100], [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
    print("%9.3f ms %6d  %5.1f%%  %s" % (a[1], a[0], 100 * a[1] / tot, k))
print("%9.3f ms total" % tot)
