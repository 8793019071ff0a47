#!/usr/bin/env python3
"""Top stalled SASS instructions of one kernel: python tools/ncu_top.py <ncu --page source --csv file> [N]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = next(r for r in rows if r and r[0] == "Address")
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows if r and r[0].startswith("0x") and len(r) > ix['stall_wait']]
S = ix['# Samples']
tot = sum(int(r[S]) for r in data)
print("total samples", tot, "instrs", len(data), "executed", sum(int(r[ix['Instructions Executed']]) for r in data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][S]))[:N]
for i in sorted(top):
    r = data[i]
    print("%5d %-66s %6s  lsb %5s wait %5s ssb %4s ex %s" % (i, r[ix['Source']][:66], r[S], r[ix['stall_long_sb']], r[ix['stall_wait']], r[ix['stall_short_sb']], r[ix['Instructions Executed']]))
if len(sys.argv) > 3:
    for a, b in [tuple(int(x) for x in s.split("-")) for s in sys.argv[3:]]:
        print("region", a, b, sum(int(r[S]) for r in data[a:b]), "of", tot)
