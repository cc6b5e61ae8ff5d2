#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, mean, min (microseconds)."""
import collections
import csv
import re
import sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    if row['Metric Name'] != 'gpu__time_duration.sum':
        continue
    k = re.sub(r'\(.*', '', row['Kernel Name']).replace('void ', '').replace('unnamed>::', '').replace('lcb::<', '')
    v = float(row['Metric Value'].replace(',', '')); u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v * 1e6 if u == 's' else v
    agg.setdefault(k, []).append(v)
print("kernel,launches,mean_us,min_us,total_us")
for k, v in agg.items():
    print(f"{k[:64]},{len(v)},{sum(v) / len(v):.1f},{min(v):.1f},{sum(v):.1f}")
