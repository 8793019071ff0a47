"""Shared helpers for the parity tests: golden-fixture access and tolerant comparison."""
import glob
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
_ALL = sorted(os.path.splitext(os.path.basename(f))[0]
              for f in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
GOLDEN = [n for n in _ALL if not n.startswith("fast_")]           # recordings of the full solver's loop
GOLDEN_FAST = [n for n in _ALL if n.startswith("fast_")]         # ... of the fast (equivalent-circuit) solver's

# Persistent state: the parity bar of BASELINE.json (1e-10 relative per step).
STATE = ["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "rho_cells", "vm_ave"]
ENV_STATE = ["E_env_x", "E_env_y", "v_env", "rho_env"]
# Diagnostics recomputed each step.  Several are sums with heavy cancellation (J_cell = Σ Jn·n̂·sa
# over a closed polygon; the divergence-free part of a gradient field in no-ECM mode), so they are
# compared against the scale of their *summands*, see scale_of().
DIAG = ["fluxes_mem", "fluxes_gj", "fluxes_env_x", "fluxes_env_y", "rate_NaKATP", "Jmem", "Jgj",
        "Jn", "I_mem", "J_cell_x", "J_cell_y", "Jc", "Eme", "E_cell_x", "E_cell_y", "Emc", "dvm",
        "J_env_x", "J_env_y", "Jtx", "Jty", "B_field", "sigma_cell", "rho_env_surf", "E_gj_x", "E_gj_y"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def group(cap, prefix):
    return {k[len(prefix):]: v for k, v in cap.items() if k.startswith(prefix)}


def mesh_of(cap, kind):
    """Mesh of a phase: the cutting event re-indexes the SIM-phase mesh (fixture default_try, oracle/refrun.py)."""
    g = group(cap, kind + ".cells.")
    return g if g else group(cap, "cells.")


def snap_steps(cap, kind):
    return sorted({int(k.split(".")[1][1:]) for k in cap if k.startswith(kind + ".k")})


def apply_schedule(obj, cap, kind, n):
    """Apply what fire_events rewrote for (1-based) step n (oracle/refrun.py)."""
    # sim modulators rewrite these inside the loop (run_loop_modulators, networks.py:3296-3299): not a schedule
    targets = [str(t) for h in range(2) for t in cap.get("%s.s0.net%d.modulator_targets" % (kind, h), [])]
    skip = {"GJ": "gj_block", "Na/K-ATPase": "NaKATP_block", "TJ": "TJ_modulator"}
    skip = {skip[t] for t in targets if t in skip}
    key = "%s.noise.k%d" % (kind, n)
    if key in cap:          # dynamic noise: the reference's own draw of this step (oracle/refrun.py)
        obj.set_noise_flux(cap[key]) if hasattr(obj, "set_noise_flux") else setattr(obj, "noise_flux", np.array(cap[key], dtype=float))
    for f, v in group(cap, "%s.sched.k%d." % (kind, n)).items():
        if f in skip:
            continue
        if f == "bound_V":
            obj.set_bound_V(v) if hasattr(obj, "set_bound_V") else setattr(
                obj, "bound_V", dict(zip("TBLR", v)))
        else:
            obj.set_field(f, v) if hasattr(obj, "set_field") else setattr(
                obj, f, np.array(v, dtype=float))


def set_time(obj, cap, kind, n):
    """Hand the time of (1-based) step n to an oracle / engine wrapper whose networks carry timed substance events."""
    if hasattr(obj, "t"):
        obj.t = float(cap[kind + ".time_steps"][n - 1])


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


EPS = 2.220446049250313e-16


def gpu_tolerances(cap, kind, ref):
    """Absolute tolerances for comparing a GPU snapshot with the reference's.

    Concentrations and gating are compared at BASELINE.json's 1e-10 relative.  Vmem, charge and
    currents are *cancelling sums* (rho = sum_i z_i F c_i is ~1e-4 of its summands; Jmem likewise;
    vgj is a difference of neighbouring Vmem), so two correct evaluation orders (BLAS dot vs an
    FMA chain) legitimately differ by the forward-error bound n*eps*sum|terms| of the sum itself;
    the reference's own np.dot is only reproducible to that bound across BLAS builds.  For those
    fields the tolerance is max(1e-10*max|ref|, 16*eps*sum|terms|) mapped through the formula.
    """
    P = group(cap, kind + ".p.")
    S0 = group(cap, kind + ".s0.")
    cells = mesh_of(cap, kind)
    zF = np.abs(np.asarray(S0["zs"], dtype=float)) * float(P["F"])
    cm, dt = float(P["cm"]), float(P["dt"])

    def mx(f):
        return float(np.max(np.abs(ref[f]))) if f in ref and np.size(ref[f]) else 0.0
    tol = {}
    for f in ("cc_cells", "cc_at_mem", "cc_env", "gjopen", "fluxes_mem", "fluxes_env_x", "fluxes_env_y",
              "rate_NaKATP", "sigma_cell"):
        tol[f] = 1e-10 * mx(f)
    b_rho = 16 * EPS * float(np.max(np.dot(zF, np.abs(ref["cc_cells"]))))
    tol["rho_cells"] = max(1e-10 * mx("rho_cells"), b_rho)
    b_vm = b_rho * float(np.max(cells["diviterm"])) / cm
    tol["vm"] = tol["vm_ave"] = max(1e-10 * mx("vm"), b_vm)
    tol["dvm"] = 2 * tol["vm"] / dt
    tol["E_gj_x"] = tol["E_gj_y"] = 4 * tol["vm"] / float(cells["gj_len"])   # -(vm[nn] - vm[m])/gj_len, sim.py:2166-2172
    if "rho_env" in ref and int(P["is_ecm"]):
        b_re = 16 * EPS * float(np.max(np.dot(zF, np.abs(ref["cc_env"]))))
        tol["rho_env"] = max(1e-10 * mx("rho_env"), b_re)
        k = tol["rho_env"] / max(mx("rho_env"), 1e-300)
        tol["v_env"] = max(1e-10 * mx("v_env"), k * mx("v_env"))
        # E = -grad(screen*v_env): differences of neighbouring v_env over delta
        if mx("v_env") > 0:
            e_scale = max(mx("E_env_x"), mx("E_env_y"))
            tol["E_env_x"] = tol["E_env_y"] = max(1e-10 * e_scale, 4 * k * e_scale * 8)
        else:
            tol["E_env_x"] = tol["E_env_y"] = 1e-300
    if "Jtx" in ref and int(P["is_ecm"]):
        # Helmholtz-Hodge parts of J_env = sum_i z_i F f_env_i (ion_current.py:50-73): the decomposition differentiates
        # and re-integrates a cancelling sum whose inputs carry 1e-10, so every part is judged against |J total|; the
        # divergence-free part J_env is 1e-6 of it on these tissues.  B_field = mu*potential: a J error times the domain size.
        sJt = max(mx("Jtx"), mx("Jty"))
        for f in ("J_env_x", "J_env_y", "Jtx", "Jty"):
            tol[f] = 1e-9 * sJt
        L = float(max(cells["grid_shape"])) * float(cells["delta"])
        tol["B_field"] = max(1e-8 * mx("B_field"), 1e-9 * sJt * float(P["mu"]) * L)
    if "Jtx" in ref and not int(P["is_ecm"]):
        # without extracellular spaces (ion_current.py:116-158): v_env is Vmem/2 scattered and smoothed, E its (Helmholtz-Hodge
        # smoothed) difference quotient times delta, J = sigma*E*D_env_weight decomposed once more
        tol["v_env"] = tol["vm"]
        e_scale = max(mx("E_env_x"), mx("E_env_y"))
        tol["E_env_x"] = tol["E_env_y"] = max(1e-9 * e_scale, 8 * tol["vm"])
        sJt = max(mx("Jtx"), mx("Jty"))
        for f in ("J_env_x", "J_env_y", "Jtx", "Jty"):
            tol[f] = 1e-9 * sJt
        L = float(max(cells["grid_shape"])) * float(cells["delta"])
        tol["B_field"] = max(1e-8 * mx("B_field"), 1e-9 * sJt * float(P["mu"]) * L)
    if "fluxes_mem" in ref:
        zs = np.asarray(S0["zs"], dtype=float)
        sJ = float(np.max(np.dot(zF, np.abs(ref["fluxes_mem"]))))
        # the GJ flux responds to vgj = vm[nn]-vm[m]: d f_gj ~ (D_gj*surf*g/len)*(zF/RT)*c*d(vgj)
        dgj = float(np.max(np.asarray(S0["D_gj"]))) * float(P["gj_surface"]) / float(cells["gj_len"])
        sens = dgj * (2 * float(P["F"]) / (float(P["R"]) * float(P["T"]))) * float(np.max(ref["cc_cells"]))
        tol["fluxes_gj"] = max(1e-10 * mx("fluxes_gj"), sens * 2 * tol["vm"])
        bJ = max(1e-10 * sJ, 16 * EPS * sJ) + float(np.max(zF)) * len(zs) * tol["fluxes_gj"]
        for f in ("Jmem", "Jgj", "Jn", "Jc", "J_cell_x", "J_cell_y"):
            tol[f] = max(1e-10 * mx(f), bJ)
        tol["I_mem"] = tol["Jn"] * float(np.max(cells["mem_sa"]))
        smin = float(np.min(ref["sigma_cell"])) if "sigma_cell" in ref else 1.0
        for f in ("E_cell_x", "E_cell_y", "Emc"):
            tol[f] = tol["Jn"] / smin
        if float(P["cell_polarizability"]) != 0.0:
            # sim.py:2059-2076: Vmem integrates Jn (dt/cm per step), and the cell field is the area-weighted sum of
            # (vm - vm_ave)/R over a closed polygon — differences of nearly equal Vmem values
            tol["vm"] = tol["vm_ave"] = max(tol["vm"], tol["Jn"] * dt / cm)
            tol["dvm"] = 2 * tol["vm"] / dt
            rmin = float(np.min(cells["R_rads"])) * float(P["true_cell_size"]) / float(P["cell_radius"])
            for f in ("E_cell_x", "E_cell_y", "Emc"):
                tol[f] = 4 * tol["vm"] / rmin
    return tol


def channels_of(cap, kind, where="s0"):
    """Channel specs recorded by tests/golden/make_golden.py:chan_extra."""
    g = group(cap, "%s.%s." % (kind, where))
    if "chan.names" not in g:
        return []
    out = []
    for k, n in enumerate(g["chan.names"]):
        pre = "chan%d." % k
        out.append({"name": str(n), "model": str(g[pre + "model"]), "ion": str(g[pre + "ion"]),
                    "maxDm": float(g[pre + "maxDm"]), "rel_perm": float(g[pre + "rel_perm"]),
                    "init_active": bool(int(g[pre + "init_active"])), "targets": g[pre + "targets"],
                    "m": g[pre + "m"], "h": g[pre + "h"]})
    return out


def fast_consts_of(cap, kind, where="s0"):
    """Constants of Simulator.fast_sim_init the fast solver's channels read (make_golden.py:chan_extra), or None."""
    g = group(cap, "%s.%s." % (kind, where))
    if "fast.rev_E" not in g:
        return None
    return {"rev_E": g["fast.rev_E"], "cbar": g["fast.cbar"], "geo_conv": float(g["fast.geo_conv"]), "zs": g["zs"]}


def network_handlers(cap, kind, where="s0"):
    """Handler ids of the recorded networks: 0 = general network (sim.molecules.core), 1 = gene regulatory network."""
    return [h for h in range(2) if "%s.%s.net%d.species" % (kind, where, h) in cap]


def networks_of(cap, kind, where="s0"):
    """Network descriptions recorded by tests/golden/make_golden.py:net_extra (general network first)."""
    from betse_b200 import network as netlib
    out = []
    for h in range(2):
        pre = "%s.%s.net%d." % (kind, where, h)
        if pre + "species" in cap:
            out.append(netlib.unflatten(cap, pre))
    return out
