#!/usr/bin/env python3
"""Bucket ncu per-instruction counters of k_mem_pipe by kernel phase.
    python tools/ncu_buckets.py <nvdisasm -g -c kmem_pipe cubin> <ncu --page source --csv> [n_tiles]"""
import sys, csv, collections, re, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_lines as nl
ins = nl.sass_lines(sys.argv[1], "k_mem_pipeILi6ELi4ELi3")
rows = list(csv.reader(open(sys.argv[2]))); hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) > 10 and r[0] != "Address"]
nt = float(sys.argv[3]) if len(sys.argv) > 3 else 199978.
src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "betse_b200/csrc/kmem_pipe.cu")).read().split("\n")
def find(s):
    for i, l in enumerate(src):
        if s in l: return i + 1
    return 10**9
L_k = find("template <int NI, int WPC, int MINB>"); L_loop = find("for (; tile < nt; tile += W, ++it)")
L_math = find("// ---- lanes = membranes"); L_slot = find("// ---- the warp's slice of the membrane->env")
L_sum = find("// ---- lanes = (cell, ion) pairs"); L_cells = find("// ---- lanes = cells: charge")
def bucket(f, l):
    if f == "kmath.cuh": return "kmath (exp/rcp/gating)"
    if f != "kmem_pipe.cu": return f
    if l < L_k: return "issue fns + waits"
    if l < L_loop: return "setup/prologue"
    if l < L_math: return "loop top: regs, issue calls"
    if l < L_slot: return "membrane math"
    if l < L_sum: return "slot copy"
    if l < L_cells: return "cell-ion sums"
    return "cells/tail"
b_ex = collections.Counter(); b_s = collections.Counter(); ops = collections.Counter()
for k in range(min(len(data), len(ins))):
    (f, l), op = ins[k]; b = bucket(f, l); n = int(data[k][ix["Instructions Executed"]])
    b_ex[b] += n; b_s[b] += int(data[k][ix["# Samples"]])
    ops[re.sub(r"^@!?U?P\d+\s+", "", op).split()[0].split(".")[0]] += n
te = sum(b_ex.values()); ts = sum(b_s.values())
print("instructions/tile %.0f" % (te / nt))
for b, c in b_s.most_common(): print("%-32s inst %5.1f%% (%6.1f/tile) samples %5.1f%%" % (b, 100 * b_ex[b] / te, b_ex[b] / nt, 100 * c / ts))
print({k: round(v / nt, 1) for k, v in ops.most_common(28)})
