#!/usr/bin/env python3
"""A/B of the drop-in loop's fixed costs on one box: python tools/e2e_ab.py [--cells N] [--steps K] [--reps R]
Alternates BETSE_PIN_PREFETCH=0/1 and prints run_sim_core_loop's own breakdown (stats['seconds'])."""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from betse_b200 import simloop, synth
from betse_b200.engine import TissueEngine

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=1_000_000); ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
mesh, p, state = synth.make_tissue(a.cells)
_m, _p, _s = synth.make_tissue(2000)
_e = TissueEngine(_m, _p, _s, device=0); _e.step(3); _e.close()
ts = np.linspace(0, a.steps * p["dt"], a.steps)
sampled = set(ts[10::10].tolist())
for rep in range(a.reps):
    for pre in ("0", "1"):
        os.environ["BETSE_PIN_PREFETCH"] = pre
        sim, phase = bench.namespaces(mesh, p, state)
        stats = {}
        t0 = time.perf_counter()
        simloop.run_sim_core_loop(sim, phase, ts, sampled, None, device=0, stats=stats)
        wall = time.perf_counter() - t0
        print("prefetch=%s wall %.3f s  %s" % (pre, wall, stats["seconds"]), flush=True)

# where the host time of one run goes, by Python function
import cProfile, pstats, io
os.environ["BETSE_PIN_PREFETCH"] = "1"
sim, phase = bench.namespaces(mesh, p, state)
pr = cProfile.Profile()
pr.enable()
simloop.run_sim_core_loop(sim, phase, ts, sampled, None, device=0, stats={})
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue())
