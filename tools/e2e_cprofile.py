#!/usr/bin/env python3
"""cProfile of the drop-in loop end to end at the driver's bench settings (1 M cells, 20 timesteps, one sample)."""
import cProfile, os, pstats, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from betse_b200 import simloop, synth
from betse_b200.engine import TissueEngine

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
mesh, p, state = synth.make_tissue(1_000_000)
_m, _p, _s = synth.make_tissue(2000)          # CUDA context + module load outside the measurements
_e = TissueEngine(_m, _p, _s, device=0); _e.step(3); _e.close()
for rep in range(2):
    sim, phase = bench.namespaces(mesh, p, state)
    ts = np.linspace(0, steps * p["dt"], steps)
    sampled = set(ts[10::10].tolist())
    stats = {}
    pr = cProfile.Profile()
    t0 = time.perf_counter()
    pr.enable()
    simloop.run_sim_core_loop(sim, phase, ts, sampled, None, device=0, stats=stats)
    pr.disable()
    print("rep", rep, "wall %.3f" % (time.perf_counter() - t0), stats.get("seconds"))
    simloop._join_closing()
pstats.Stats(pr).sort_stats("tottime").print_stats(28)
