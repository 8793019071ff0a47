#!/usr/bin/env python3
"""Attribute ncu per-SASS-instruction counters to CUDA source lines.

    python tools/ncu_lines.py <nvdisasm -g -c output> <mangled-name substring> <ncu --page source --csv> [top N]

The ncu CSV lists the kernel's SASS in address order; nvdisasm -g interleaves the same SASS with
`//## File "...", line N` markers.  The two are zipped by instruction index."""
import collections
import csv
import re
import sys


def sass_lines(fn, func):
    out = []
    active = False
    cur = ("?", 0)
    for ln in open(fn):
        if ln.startswith("\t.section\t.text."):
            active = func in ln
            continue
        if ln.startswith("\t.section") or ln.startswith("//-----"):
            if active and out and ".text." not in ln:
                pass
            if ln.startswith("\t.section"):
                active = False
            continue
        if not active:
            continue
        m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            out.append((cur, m.group(2)))
    return out


def main():
    sass, func, ncsv = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    ins = sass_lines(sass, func)
    rows = list(csv.reader(open(ncsv)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) > 10 and r[0] != "Address"]
    if len(data) != len(ins):
        print("warning: %d ncu rows vs %d sass instructions" % (len(data), len(ins)))
    n = min(len(data), len(ins))
    ex = collections.Counter()
    smp = collections.Counter()
    for k in range(n):
        (f, l), _ = ins[k]
        ex[(f, l)] += int(data[k][ix["Instructions Executed"]])
        smp[(f, l)] += int(data[k][ix["# Samples"]])
    te, ts = sum(ex.values()), sum(smp.values())
    print("total executed %d, samples %d" % (te, ts))
    src = {}
    for (f, l), c in sorted(smp.items(), key=lambda kv: -kv[1])[:top]:
        if f not in src:
            try:
                src[f] = open("betse_b200/csrc/" + f).read().split("\n")
            except Exception:
                src[f] = []
        text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
        print("%-22s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (f, l, 100.0 * ex[(f, l)] / te, 100.0 * c / ts, text))


if __name__ == "__main__":
    main()
