"""Print the worst GPU-vs-reference error per field and snapshot for every golden fixture."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util
from betse_b200.engine import TissueEngine

out = open("gpurun_out/report.txt", "w")
def P(*a):
    s = " ".join(str(x) for x in a); print(s); out.write(s + "\n")
for name in util.GOLDEN:
    if "polar" in name or "chan" in name or "net" in name: continue
    cap = util.load_golden(name)
    for kind in ("init", "sim"):
        eng = TissueEngine(util.group(cap, "cells."), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
        ecm = bool(int(cap[kind + ".p.is_ecm"]))
        snaps = util.snap_steps(cap, kind); n = 0
        for K in snaps:
            last = K == snaps[-1]
            while n < K:
                util.apply_schedule(eng, cap, kind, n + 1)
                st = eng.step(1, diag=(last and n + 1 == K)); n += 1
            ref = util.group(cap, "%s.k%d." % (kind, K))
            fields = list(util.STATE) + (util.ENV_STATE if ecm else [])
            if last:
                fields += [f for f in util.DIAG if f in ref and f not in ("rho_env_surf","Eme") and (ecm or f not in ("J_env_x","J_env_y","Jtx","Jty","B_field"))]
                if not ecm: fields = [f for f in fields if not f.startswith("fluxes_env")]
            got = eng.download([f for f in fields if f in ref])
            errs = {f: util.rel_err(a, ref[f], util.scale_of(f, ref)) for f, a in got.items()}
            P(name, kind, "k=%d" % K, "status", st, " ".join("%s=%.1e" % (f, e) for f, e in errs.items()))
        eng.close()
