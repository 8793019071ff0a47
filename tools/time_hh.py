#!/usr/bin/env python3
"""Sampled (diagnostics) step of the 1 M-cell tissue: wall time per step with / without diagnostics.  Under
`ncu --metrics gpu__time_duration.sum -k regex:k_hh` the launch list gives the Helmholtz-Hodge kernels themselves."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from betse_b200 import synth
from betse_b200.engine import TissueEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
mesh, p, st = synth.make_tissue(n)
eng = TissueEngine(mesh, p, st)
eng.update_V()
eng.step(3)
eng.step(1, diag=True)
for diag in (False, True):
    t0 = time.perf_counter()
    for _ in range(5):
        eng.step(1, diag=diag)
    print("diag" if diag else "plain", "%.3f ms/step (wall, one step per call)" % ((time.perf_counter() - t0) / 5 * 1e3))
got = eng.download(["J_env_x", "B_field"])
print("finite", all(np.isfinite(a).all() for a in got.values()), float(np.abs(got["J_env_x"]).max()))
eng.close()
