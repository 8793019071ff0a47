#!/usr/bin/env python3
"""Kernels of the last timestep of an ncu launch list: python tools/launch_step.py <launches.csv>  (time, DRAM bytes per launch)"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
ix = {h: i for i, h in enumerate(rows[0])}
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(int(r[ix['ID']]), {'name': r[ix['Kernel Name']]})
    v = float(r[ix['Metric Value']].replace(',', ''))
    u = r[ix['Metric Unit']]
    v *= {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1.0)
    d[r[ix['Metric Name']]] = v
ids = list(per)
ends = [i for i, k in enumerate(ids) if 'k_field' in per[k]['name']]
a, b = (ends[-2] + 1, ends[-1] + 1) if len(ends) > 1 else (0, len(ids))
tot = 0.0
for k in ids[a:b]:
    d = per[k]
    tot += d['gpu__time_duration.sum']
    print("  %-30s %8.2f us  rd %7.2f MB  wr %7.2f MB" % (d['name'][:30], d['gpu__time_duration.sum'], d.get('dram__bytes_read.sum', 0) / 1e6, d.get('dram__bytes_write.sum', 0) / 1e6))
print("  %d kernels, %.1f us" % (b - a, tot))
