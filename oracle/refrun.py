"""Test infrastructure ONLY: drive the REAL reference (``/root/reference``) end to end —
``SimRunner.seed`` → ``init`` → ``sim`` — in this container and record what its
``Simulator._run_sim_core_loop`` (betse/science/sim.py:1132-1390) sees and produces.

The recording is a flat ``{name: ndarray}`` dict ("capture") that
``tests/golden/make_golden.py`` commits as ``tests/golden/*.npz``:

* ``cells.*``   – the mesh / index arrays the loop consumes (cells.py, SURVEY §2 ★data)
* ``p.*``       – every scalar parameter the loop reads
* ``<phase>.s0.*`` – ``Simulator`` state on entry to the loop (after ``init_dynamics`` +
  first ``update_V``, sim.py:1034-1041)
* ``<phase>.k<N>.*`` – state and diagnostics after N timesteps of the reference loop
* ``<phase>.sched.k<N>.*`` – value of anything ``fire_events`` rewrote, in effect DURING step N

The reference loop is advanced by calling the *unmodified* method on consecutive
single-step slices of ``time_steps`` – the loop body carries no cross-iteration locals, so
this is identical to one call (checked in tests/test_oracle_vs_reference.py).

Not importable on the GPU box (no reference there) and never imported by the product.
"""
import copy
import os
import shutil
import tempfile

import numpy as np

from . import refshim

# Simulator attributes recorded at loop entry (persistent state + frozen constants).
STATE_FIELDS = [
    "cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "Dm_cells", "D_gj", "D_free", "zs",
    "c_env_bound", "T", "vgj",
    "extra_rho_cells", "extra_rho_env", "extra_J_mem", "extra_Jenv_x", "extra_Jenv_y",
    "smooth_weight_mem", "smooth_weight_o", "ko_env", "rho_factor", "sigma",
    "NaKATP_block", "gj_block", "CaATP_block", "rho_pump", "rho_channel",
    "D_env", "TJ_modulator", "E_env_x", "E_env_y", "v_env", "rho_env", "Phi_b", "D_env_weight",
    "rho_cells", "vm_ave", "Jn", "envV",
    "E_cell_x", "E_cell_y",        # read by Molecule.update_intra of charged substances in the first step
]
# The fast (equivalent-circuit) solver's own state and constants (Simulator.fast_sim_init, sim.py:1393-1452; loop sim.py:1454-1640)
FAST_FIELDS = ["G_Leak", "E_Leak", "G_gj", "vm_GHK", "Emx", "Emy", "sigma_cell", "J_cell_x", "J_cell_y"]
# Additional per-step outputs (diagnostics recomputed from scratch every step).
DIAG_FIELDS = [
    "fluxes_mem", "fluxes_gj", "fluxes_env_x", "fluxes_env_y", "rate_NaKATP", "rate_CaATP",
    "Jmem", "Jgj", "I_mem", "J_cell_x", "J_cell_y", "Jc", "J_env_x", "J_env_y", "Jtx", "Jty",
    "B_field", "Eme", "E_cell_x", "E_cell_y", "Emc", "dvm", "Egj", "E_gj_x", "E_gj_y",
    "sigma_cell", "rho_env_surf",
]
CELLS_FIELDS = [
    "mem_to_cells", "nn_i", "bflags_mems", "map_mem2ecm", "map_cell2ecm", "mem_sa", "R_rads",
    "cell_vol", "cell_sa", "diviterm", "num_mems", "delta", "memSa_per_envSquare", "ecm_vol",
    "gj_default_weights", "cell_centres", "mem_mids_flat", "xypts", "envInds_inClust",
    "all_bound_mem_inds", "interior_bound_mem_inds", "ecm_inds_bound_cell", "bflags_cells",
    "mem_vol", "true_ecm_vol",
]
P_FIELDS = [
    "F", "R", "T", "q", "kb", "eo", "er", "cm", "tm", "dt", "NAv", "mu",
    "alpha_NaK", "alpha_Ca", "KmNK_Na", "KmNK_K", "KmNK_ATP", "KmCa_Ca", "KmCa_ATP",
    "cATP", "cADP", "cPi", "deltaGATP", "gj_surface", "gj_vthresh", "gj_min",
    "v_sensitive_gj", "cluster_open", "is_ecm", "vol_env", "cell_height", "cell_space",
    "fast_update_ecm", "sharpness", "cell_radius", "true_cell_size", "smooth_cells",
    "cell_polarizability", "D_tj", "D_adh", "channel_noise_level", "dynamic_noise",
    "dynamic_noise_level", "init_tsteps", "sim_tsteps", "t_resample", "Ca_dyn",
    "molecules_enabled", "grn_enabled", "fluid_flow", "deform_osmo", "deformation",
    "substances_affect_charge",
]


def _arr(v):
    if isinstance(v, dict):
        return None
    try:
        a = np.asarray(v)
    except Exception:
        return None
    if a.dtype == object:
        return None
    return a.copy()


def snapshot(sim, fields):
    out = {}
    for f in fields:
        if hasattr(sim, f):
            a = _arr(getattr(sim, f))
            if a is not None:
                out[f] = a
    if hasattr(sim, "bound_V"):
        bv = sim.bound_V
        out["bound_V"] = np.array([bv["T"], bv["B"], bv["L"], bv["R"]], dtype=float)
    return out


def snapshot_cells(cells):
    out = {}
    for f in CELLS_FIELDS:
        if hasattr(cells, f):
            a = _arr(getattr(cells, f))
            if a is not None:
                out[f] = a
    out["mem_nx"] = cells.mem_vects_flat[:, 2].copy()
    out["mem_ny"] = cells.mem_vects_flat[:, 3].copy()
    out["grid_shape"] = np.array(cells.X.shape, dtype=np.int64)
    out["gj_len"] = np.asarray(float(cells.gj_len))
    # cell_to_mems is a ragged object array; the loop only relies on it being the CSR of
    # mem_to_cells (membranes of a cell contiguous, cells.py:1095-1146) – record the pointer.
    ptr = np.zeros(len(cells.cell_i) + 1, dtype=np.int64)
    for ci, mems in enumerate(cells.cell_to_mems):
        mems = np.asarray(mems)
        assert mems.size and np.all(np.diff(mems) == 1) and mems[0] == ptr[ci], "non-contiguous"
        ptr[ci + 1] = mems[-1] + 1
    out["cell_mem_ptr"] = ptr
    return out


def snapshot_p(p):
    out = {}
    for f in P_FIELDS:
        if hasattr(p, f):
            v = getattr(p, f)
            if isinstance(v, (bool, np.bool_)):
                v = int(v)
            if v is None:
                continue
            out[f] = np.asarray(v)
    out["ions"] = np.array([k for k, v in p.ions_dict.items() if v == 1])
    return out


def write_config(dst_dir, mods):
    """Copy the reference's shipped config tree and apply ``mods`` (nested dict) to it."""
    import yaml as pyyaml
    import ruamel.yaml as ry  # the refshim stand-in (YAML 1.2 resolver on PyYAML)
    src = os.path.join(refshim.REF_ROOT, "betse", "data", "yaml")
    shutil.copytree(src, dst_dir, dirs_exist_ok=True)
    for root, dirs, files in os.walk(dst_dir):
        for n in dirs + files:
            os.chmod(os.path.join(root, n), 0o755)
    fn = os.path.join(dst_dir, "sim_config.yaml")
    with open(fn) as f:
        conf = ry.YAML().load(f)

    def merge(d, m):
        for k, v in m.items():
            if isinstance(v, dict) and isinstance(d.get(k), dict):
                merge(d[k], v)
            else:
                d[k] = v
    merge(conf, mods)
    with open(fn, "w") as f:
        pyyaml.safe_dump(conf, f, default_flow_style=False, sort_keys=False)
    return fn


class LoopRecorder:
    """Wraps the reference's Simulator._run_sim_core_loop to record entry state and the
    state after chosen step counts, without altering what the reference computes."""

    def __init__(self, snap_steps, max_steps=None, extra=None, trace=(), precut=False, method="_run_sim_core_loop"):
        self.method = method        # "_run_fast_sim_core_loop": the equivalent-circuit solver (sim.py:1454-1640)
        self.snap_steps = {k: sorted(set(v)) for k, v in snap_steps.items()}
        self.max_steps = max_steps or {}
        self.capture = {}
        self.extra = extra  # optional callable(sim, phase) -> dict of extra arrays
        self.trace = tuple(trace)   # fields recorded at EVERY sampled step: <phase>.trace.<field> [n_sampled, ...]
        self.precut = precut        # fire a pending cutting event before the entry snapshot (see wrapped)

    def install(self):
        from betse.science.sim import Simulator
        rec = self
        orig = getattr(Simulator, self.method)
        self._orig = orig
        more = FAST_FIELDS if self.method == "_run_fast_sim_core_loop" else []

        def wrapped(sim, phase, time_steps, time_steps_sampled, anim_cells):
            kind = phase.kind.name.lower()
            cap = rec.capture
            dyna = getattr(phase, "dyna", None)
            ev = getattr(dyna, "event_cut", None)
            if rec.precut and kind == "sim" and ev is not None and not ev.is_fired and len(time_steps):
                # The cutting event fires inside the FIRST fire_events call of the phase (tishandler.py:884-913,
                # event_cut_time == 0, parameters.py:687) and re-indexes every array.  Firing it here, before the
                # entry snapshot, is the same computation (fire_events is the first statement of the loop body that
                # touches state, sim.py:1187-1188, and is a pure function of t once the cut has fired); the fixture
                # then holds the post-cut mesh (sim.cells.*) and state.
                dyna.fire_events(phase=phase, t=time_steps[0])
                for k, v in snapshot_cells(phase.cells).items():
                    cap["sim.cells." + k] = v
            if "cells.mem_sa" not in cap:
                for k, v in snapshot_cells(phase.cells).items():
                    cap["cells." + k] = v
            for k, v in snapshot_p(phase.p).items():
                cap["%s.p.%s" % (kind, k)] = v
            for k, v in snapshot(sim, STATE_FIELDS + more).items():
                cap["%s.s0.%s" % (kind, k)] = v
            if rec.extra:
                for k, v in rec.extra(sim, phase).items():
                    cap["%s.s0.%s" % (kind, k)] = v
            cap[kind + ".time_steps"] = np.asarray(time_steps).copy()
            cap[kind + ".time_steps_sampled"] = np.array(sorted(time_steps_sampled))
            snaps = rec.snap_steps.get(kind, [])
            nmax = rec.max_steps.get(kind, len(time_steps))
            # Host-side schedule (tissue/tishandler.py:709-917, 1321-1332): everything
            # fire_events may rewrite, recorded whenever the value in effect DURING step n
            # (1-based) differs from the one in effect during step n-1 / at loop entry.
            sched_fields = ("Dm_cells", "D_env", "TJ_modulator", "gj_block", "NaKATP_block",
                            "c_env_bound", "T", "bound_V", "D_gj")
            prev = snapshot(sim, sched_fields)
            tr = {f: [] for f in rec.trace}
            tr_steps = []
            for n in range(min(nmax, len(time_steps))):
                orig(sim, phase=phase, time_steps=time_steps[n:n + 1],
                     time_steps_sampled=time_steps_sampled, anim_cells=anim_cells)
                if kind == "sim" and getattr(phase.p, "dynamic_noise", False) and hasattr(sim, "protein_noise_flux"):
                    # the random walk on the protein concentration of THIS step (sim.py:1322-1339): the replay needs the draw
                    cap["sim.noise.k%d" % (n + 1)] = np.array(sim.protein_noise_flux, dtype=float, copy=True)
                cur = snapshot(sim, sched_fields)
                for f, v in cur.items():
                    if f not in prev or prev[f].shape != v.shape or not np.array_equal(prev[f], v):
                        cap["%s.sched.k%d.%s" % (kind, n + 1, f)] = v
                prev = cur
                if rec.trace and time_steps[n] in time_steps_sampled:
                    tr_steps.append(n + 1)
                    for f in rec.trace:
                        tr[f].append(np.array(getattr(sim, f), dtype=float, copy=True))
                if (n + 1) in snaps:
                    for k, v in snapshot(sim, STATE_FIELDS + DIAG_FIELDS + more).items():
                        cap["%s.k%d.%s" % (kind, n + 1, k)] = v
                    if rec.extra:
                        for k, v in rec.extra(sim, phase).items():
                            cap["%s.k%d.%s" % (kind, n + 1, k)] = v

            if rec.trace:
                cap[kind + ".trace.steps"] = np.array(tr_steps, dtype=np.int64)
                for f in rec.trace:
                    cap["%s.trace.%s" % (kind, f)] = np.array(tr[f])

        setattr(Simulator, self.method, wrapped)
        return self

    def uninstall(self):
        from betse.science.sim import Simulator
        setattr(Simulator, self.method, self._orig)


def run_reference(mods, seed=12345, snap_steps=None, max_steps=None, phases=("init", "sim"),
                  extra=None, workdir=None, keep=False, tweak_p=None, trace=(), precut=False, method="_run_sim_core_loop"):
    """Run the real reference on the shipped default config + ``mods``.

    Returns the capture dict.  ``np.random.seed(seed)`` is set once before ``seed`` (the
    reference never seeds: SURVEY §4) so mesh + channel noise are reproducible.
    """
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks

    snap_steps = snap_steps or {"init": [1, 2, 5], "sim": [1, 2, 5]}
    tmp = workdir or tempfile.mkdtemp(prefix="betse_ref_")
    try:
        fn = write_config(tmp, mods)
        np.random.seed(seed)
        p = Parameters.make(fn)
        p.anim.is_while_sim = False
        p.anim.is_after_sim = False
        p.plot.is_after_sim = False
        if tweak_p:
            tweak_p(p)
        rec = LoopRecorder(snap_steps, max_steps=max_steps, extra=extra, trace=trace, precut=precut, method=method).install()
        try:
            runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
            runner.seed()
            if "init" in phases:
                runner.init()
            if "sim" in phases:
                runner.sim()
        finally:
            rec.uninstall()
        return rec.capture
    finally:
        if not keep and workdir is None:
            shutil.rmtree(tmp, ignore_errors=True)
