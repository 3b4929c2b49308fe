#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total time, share.
usage: python tools/launch_summary.py launches.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit() and r[12] == 'gpu__time_duration.sum']
agg = {}
for r in rows:
    a = agg.setdefault(r[4], [0, 0.0]); a[0] += 1; a[1] += float(r[14].replace(',', '')) * {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(r[13], 1)
tot = sum(a[1] for a in agg.values())
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'{k[:90]:90s} launches {a[0]:3d} total_ns {a[1]:12.0f} share {100 * a[1] / tot:5.1f}%')
