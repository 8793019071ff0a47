#!/usr/bin/env python3
"""Where the end-to-end time of the drop-in loop goes (1 M cells): engine creation, state upload, timesteps,
sampled-step downloads.  python tools/e2e_profile.py [--cells N]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from betse_b200 import simloop, synth
from betse_b200.engine import TissueEngine

ap = argparse.ArgumentParser(); ap.add_argument("--cells", type=int, default=1_000_000); a = ap.parse_args()
mesh, p, state = synth.make_tissue(a.cells)
sim, phase = bench.namespaces(mesh, p, state)
T = {}
def tic(): return time.perf_counter()
_m, _p, _s = synth.make_tissue(2000)          # CUDA context + module load outside the measurements
_e = TissueEngine(_m, _p, _s, device=0); _e.step(3); _e.close()
t = tic(); m = simloop.mesh_from_cells(phase.cells); pr = simloop.params_from_p(phase.p); st = simloop.state_from_sim(sim); T["shim dicts"] = tic() - t
t = tic(); eng = TissueEngine(m, pr, st, device=0); T["TissueEngine() incl. upload"] = tic() - t
t = tic(); eng.step(10, diag=True); T["10 steps (first: graph build)"] = tic() - t
t = tic(); eng.step(10, diag=True); T["10 steps"] = tic() - t
for grp, fields in (("state", simloop._SAMPLED_STATE), ("env", simloop._SAMPLED_ENV), ("diag", simloop._SAMPLED_DIAG), ("diag_env", simloop._SAMPLED_DIAG_ENV)):
    t = tic(); got = eng.download(list(fields)); dt = tic() - t
    nb = sum(x.nbytes for x in got.values())
    T["download %s (%d MB)" % (grp, nb >> 20)] = dt
t = tic(); simloop._copy_back(sim, eng, diag=True); T["_copy_back(diag=True)"] = tic() - t
t = tic(); got = eng.download(list(simloop._W2S_STATE + simloop._W2S_ENV + simloop._W2S_DIAG + ["J_env_x", "J_env_y"]), pinned=True); T["download write2storage set, pinned, first (alloc)"] = tic() - t
t = tic(); got = eng.download(list(simloop._W2S_STATE + simloop._W2S_ENV + simloop._W2S_DIAG + ["J_env_x", "J_env_y"]), pinned=True); T["download write2storage set, pinned (%d MB)" % (sum(x.nbytes for x in got.values()) >> 20)] = tic() - t
t = tic(); eng.step(1, diag=True); T["1 step with diagnostics (k_diag + Helmholtz-Hodge)"] = tic() - t
t = tic(); eng.step(1); T["1 step"] = tic() - t
t = tic()
eng.__dict__.pop("_pinned", {}).clear()
T["close: free pinned staging"] = tic() - t
t = tic(); eng.close(); T["close: betse_destroy"] = tic() - t
for k, v in T.items(): print("%-40s %8.1f ms" % (k, v * 1e3))
