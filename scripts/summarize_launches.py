#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if "Kernel Name" in r:
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        u = d.get("Metric Unit", "ns")
        ns = v * (1e3 if u.startswith("us") else 1e6 if u.startswith("ms") else 1e9 if u == "s" else 1.0)
        k = d["Kernel Name"][:80]
        agg[k][0] += 1
        agg[k][1] += ns
tot = sum(v[1] for v in agg.values()) or 1.0
print("total kernel time %.3f ms over %d launches" % (tot / 1e6, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-82s n=%6d total=%10.1f us avg=%9.2f us share=%5.1f%%" % (k, v[0], v[1] / 1e3, v[1] / 1e3 / v[0], 100 * v[1] / tot))
