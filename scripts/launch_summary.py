#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
usage: python scripts/launch_summary.py profiles/launches_rNN.csv"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, per = 0.0, defaultdict(lambda: [0.0, 0])
for r in rows[1:]:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}.get(r[mu], 1e-3)
    per[r[kn]][0] += v * scale
    per[r[kn]][1] += 1
    tot += v * scale
for k, (t, n) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    print(f"{t / tot * 100:6.2f}% {t / 1e3:10.3f} ms  n={n:3d} avg={t / n:9.1f} us  {k[:110]}")
