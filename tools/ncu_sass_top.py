#!/usr/bin/env python3
"""Top stalled SASS instructions of a kernel in an .ncu-rep:  python tools/ncu_sass_top.py <rep> <kernel regex> [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "-k", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "hdr": None, "data": []}
        sections.append(cur)
    elif cur is not None and r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and cur["hdr"] and len(r) == len(cur["hdr"]):
        cur["data"].append(r)
for sec in sections[:1]:
    hdr, data = sec["hdr"], sec["data"]
    ix = {h: i for i, h in enumerate(hdr)}
    tot = sum(int(r[ix["# Samples"]]) for r in data)
    execd = sum(int(r[ix["Instructions Executed"]]) for r in data)
    print(sec["name"], "samples", tot, "SASS", len(data), "warp-instr executed", execd)
    keys = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(r[ix[k]]) for r in data) for k in keys}
    print("stall mix:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / tot) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
    for j, r in sorted(enumerate(data), key=lambda x: -int(x[1][ix["# Samples"]]))[:N]:
        s = int(r[ix["# Samples"]])
        parts = {k: int(r[ix[k]]) for k in keys}
        main = max(parts, key=parts.get)
        print("%5.2f%% #%-5d %-72s %s=%d exec=%s" % (100 * s / tot, j, r[ix["Source"]].strip()[:72], main[6:], parts[main], r[ix["Instructions Executed"]]))
