"""Shared helpers for the parity tests: golden-fixture access and tolerant comparison."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN = sorted(os.path.splitext(os.path.basename(f))[0]
                for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))

# Persistent state: the parity bar of BASELINE.json (1e-10 relative per step).
STATE = ["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "rho_cells", "vm_ave"]
ENV_STATE = ["E_env_x", "E_env_y", "v_env", "rho_env"]
# Diagnostics recomputed each step.  Several are sums with heavy cancellation (J_cell = Σ Jn·n̂·sa
# over a closed polygon; the divergence-free part of a gradient field in no-ECM mode), so they are
# compared against the scale of their *summands*, see scale_of().
DIAG = ["fluxes_mem", "fluxes_gj", "fluxes_env_x", "fluxes_env_y", "rate_NaKATP", "Jmem", "Jgj",
        "Jn", "I_mem", "J_cell_x", "J_cell_y", "Jc", "Eme", "E_cell_x", "E_cell_y", "Emc", "dvm",
        "J_env_x", "J_env_y", "Jtx", "Jty", "B_field", "sigma_cell", "rho_env_surf"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def group(cap, prefix):
    return {k[len(prefix):]: v for k, v in cap.items() if k.startswith(prefix)}


def snap_steps(cap, kind):
    return sorted({int(k.split(".")[1][1:]) for k in cap if k.startswith(kind + ".k")})


def apply_schedule(obj, cap, kind, n):
    """Apply what fire_events rewrote for (1-based) step n (oracle/refrun.py)."""
    for f, v in group(cap, "%s.sched.k%d." % (kind, n)).items():
        if f == "bound_V":
            obj.set_bound_V(v) if hasattr(obj, "set_bound_V") else setattr(
                obj, "bound_V", dict(zip("TBLR", v)))
        else:
            obj.set_field(f, v) if hasattr(obj, "set_field") else setattr(
                obj, f, np.array(v, dtype=float))


def scale_of(field, ref):
    """Magnitude against which the error of ``field`` is judged."""
    def mx(f):
        return float(np.max(np.abs(ref[f]))) if f in ref and np.size(ref[f]) else 0.0
    if field in ("J_cell_x", "J_cell_y", "Jc"):
        return max(mx("Jn"), mx(field))
    if field in ("E_cell_x", "E_cell_y", "Emc"):
        s = mx("sigma_cell")
        return max(mx("Jn") / s if s else 0.0, mx(field))
    if field in ("J_env_x", "J_env_y", "B_field"):
        return max(mx("Jtx"), mx("Jty"), mx(field)) * (1.0 if field != "B_field" else 1.0)
    if field == "dvm":
        return max(mx("vm") / 1e-4, mx(field))
    return mx(field)


def rel_err(a, r, scale=None):
    a = np.asarray(a, dtype=float).reshape(np.shape(r))
    r = np.asarray(r, dtype=float)
    if r.size == 0:
        return 0.0
    s = scale if scale else float(np.max(np.abs(r)))
    e = float(np.max(np.abs(a - r)))
    return e / s if s > 0 else e
