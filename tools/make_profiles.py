"""Extract the ion-profile / parameter defaults the synthetic-tissue generator needs from the
golden captures (i.e. from the real reference's Parameters + Simulator.init_core output) into
betse_b200/data/profiles.json.  Sources: parameters.py:1064-1148 (constants),
parameters.py:1265-1500 (_load_ion_profile), data/yaml/sim_config.yaml (tissue 'Base' Dm)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import util

out = {}
for prof, fx in (("basic", "basic_ecm"), ("mammal", "mammal_ecm")):
    cap = util.load_golden(fx)
    P = util.group(cap, "init.p.")
    S = util.group(cap, "init.s0.")
    d = {"ions": [str(x) for x in P["ions"]],
         "zs": [float(x) for x in S["zs"]],
         "D_free": [float(x) for x in S["D_free"]],
         "cell_concs": [float(x) for x in S["cc_cells"][:, 0]],
         "env_concs": [float(x) for x in S["cc_env"][:, 0]],
         "Dm_base": [float(x) for x in S["Dm_cells"][:, 0]],
         "p": {k: (float(v) if np.asarray(v).dtype.kind == "f" else int(v))
               for k, v in P.items() if k != "ions" and np.asarray(v).ndim == 0}}
    d["p"].pop("init_tsteps", None); d["p"].pop("sim_tsteps", None)
    out[prof] = d
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "betse_b200", "data", "profiles.json"), "w"), indent=1)
print(json.dumps(out["mammal"], indent=1)[:1500])
