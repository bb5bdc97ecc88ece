#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel.
usage: summarise_launches.py <launches.csv> "<header comment>" > out.csv"""
import csv, sys, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]; ik, iv = h.index("Kernel Name"), h.index("Metric Value")
n = collections.Counter(); t = collections.Counter()
for r in rows[1:]:
    k = r[ik][:70]; n[k] += 1; t[k] += float(r[iv].replace(",", ""))
tot = sum(t.values())
print("# " + (sys.argv[2] if len(sys.argv) > 2 else ""))
print("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 ; times are cold-cache, serialised (shares, not absolutes)")
print("kernel,launches,total_ns,share")
for k, v in t.most_common():
    print('"%s",%d,%d,%.3f' % (k, n[k], v, v / tot))
