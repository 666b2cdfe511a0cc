#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if r[0] == "ID":
        hdr = r
        continue
    if hdr is None:
        continue
    name = r[hdr.index("Kernel Name")]
    val = float(r[hdr.index("Metric Value")].replace(",", ""))
    unit = r[hdr.index("Metric Unit")]
    val = val / 1e3 if unit == "ns" else val * 1e3 if unit == "ms" else val
    agg[name[:64]][0] += 1
    agg[name[:64]][1] += val
tot = sum(v[1] for v in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]:10.1f} us {v[0]:4d} launches {v[1] / v[0]:8.1f} us/launch {100 * v[1] / tot:5.1f}%  {k}")
