"""Drop-in replacement for ``Simulator._run_sim_core_loop`` (betse/science/sim.py:1132-1390).

The reference picks its time loop through a bound-method dispatch
(``solver_method = self._run_sim_core_loop``, sim.py:1064-1075).  ``install()`` rebinds that
method to :func:`run_sim_core_loop`, which keeps the reference's contract —

* on entry ``init_dynamics``, ``clear_storage`` and ``update_V`` have run (sim.py:1034-1041), so
  every ``Simulator`` array exists as host NumPy;
* SIM phase: ``phase.dyna.fire_events(phase, t)`` is called every step (sim.py:1187-1188) — it is
  host-side scalar scheduling and stays the reference's own code; whatever it rewrote is pushed
  to the device before the step;
* at every sampled step the attributes ``write2storage`` reads (sim.py:1796-1877) hold that step's
  values as host NumPy, ``phase.callbacks.progressed_next()`` is called exactly once, then
  ``write2storage`` and the optional animation frame (sim.py:1368-1381);
* NaN in Vmem / a concentration raises ``BetseSimUnstableException`` (sim.py:1363,
  sim_toolbox.py:332-344, 475-503) after the partial state has been copied back, so that
  ``run_sim_core`` can still pickle it (sim.py:1104-1128)

— while the per-timestep arithmetic runs in libbetse_b200.so with the state resident in HBM.
The function is duck-typed on ``sim`` / ``phase.cells`` / ``phase.p`` (no reference import),
so the same code path is exercised on the GPU box from recorded reference states.
"""
import os
import time

import numpy as np

from . import capi
from .capi import BetseB200Error
from .engine import PinnedPrefetch, TissueEngine, _DOWN_SHAPES
from .network import event_values as netlib_event_values

_P_FIELDS = [
    "F", "R", "T", "q", "kb", "eo", "er", "cm", "tm", "dt", "NAv", "mu", "alpha_NaK", "alpha_Ca",
    "KmNK_Na", "KmNK_K", "KmNK_ATP", "KmCa_Ca", "KmCa_ATP", "cATP", "cADP", "cPi", "deltaGATP",
    "gj_surface", "gj_vthresh", "gj_min", "v_sensitive_gj", "cluster_open", "is_ecm", "vol_env",
    "cell_height", "cell_space", "fast_update_ecm", "sharpness", "cell_radius", "true_cell_size",
    "smooth_cells", "cell_polarizability", "substances_affect_charge",
]
_STATE_FIELDS = [
    "cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "Dm_cells", "D_gj", "D_free", "zs",
    "c_env_bound", "T", "extra_rho_cells", "extra_rho_env", "extra_J_mem", "ko_env",
    "NaKATP_block", "gj_block", "rho_pump", "rho_channel", "D_env", "TJ_modulator", "E_env_x",
    "E_env_y", "Phi_b", "D_env_weight", "sigma", "E_cell_x", "E_cell_y",
]
# what fire_events / makeAllChanges may rewrite between steps (tishandler.py:728-917, 1321-1332)
_SCHEDULED = ["Dm_cells", "D_env", "TJ_modulator", "gj_block", "NaKATP_block", "c_env_bound", "T"]



def _live_scheduled(p):
    """The scheduled quantities this phase's events CAN move (tishandler.py:744-876 tests the same option entries before
    it touches anything): comparing the others after every fire_events call would cost a pass over arrays of 10^7 entries
    per time step for nothing (makeAllChanges rebinds sim.Dm_cells to an equal array each step, tishandler.py:1327-1330).
    Unknown parameter objects: everything is live."""
    go, so = getattr(p, "global_options", None), getattr(p, "scheduled_options", None)
    if not isinstance(go, dict) or not isinstance(so, dict):
        return list(_SCHEDULED)

    def on(d, *keys):
        for k in keys:
            v = d.get(k, 0)
            if isinstance(v, (list, tuple, dict, np.ndarray)) or v is None or v != 0:      # the reference stores 0 for "off"
                return True
        return False
    live = []
    if on(so, "Na_mem", "K_mem", "Cl_mem", "Ca_mem"):
        live.append("Dm_cells")
    if on(so, "ecmJ"):
        live.append("D_env")
    if on(go, "gj_block"):
        live.append("gj_block")
    if on(go, "NaKATP_block"):
        live.append("NaKATP_block")
    if on(go, "K_env", "Cl_env", "Na_env"):
        live.append("c_env_bound")
    if on(go, "T_change"):
        live.append("T")
    return live


# attributes refreshed at sampled steps (read by write2storage and the exporters)
_SAMPLED_STATE = ["cc_cells", "cc_at_mem", "cc_env", "vm", "vm_ave", "gjopen", "rho_cells", "Phi_b"]
_SAMPLED_ENV = ["E_env_x", "E_env_y", "v_env", "rho_env"]
_SAMPLED_DIAG = ["fluxes_mem", "fluxes_gj", "rate_NaKATP", "Jmem", "Jgj", "Jn", "I_mem", "Jc", "Emc",
                 "dvm", "J_cell_x", "J_cell_y", "E_cell_x", "E_cell_y", "sigma_cell"]
_SAMPLED_DIAG_ENV = ["fluxes_env_x", "fluxes_env_y", "J_env_x", "J_env_y", "B_field", "Jtx", "Jty"]
# exactly what Simulator.write2storage reads (sim.py:1796-1877) — the per-sample download; everything else
# (per-ion flux arrays, J's, E_cell ...) is only read after the phase and comes with the final copy-back
_W2S_STATE = ["cc_cells", "vm", "vm_ave", "gjopen", "rho_cells"]
_W2S_ENV = ["cc_env", "E_env_x", "E_env_y", "v_env"]
_W2S_DIAG = ["I_mem", "J_cell_x", "J_cell_y", "rate_NaKATP", "E_gj_x", "E_gj_y"]


class SimUnstable(Exception):
    """Raised when the reference's BetseSimUnstableException is not importable."""


def _unstable_exception():
    try:
        from betse.exceptions import BetseSimUnstableException
        return BetseSimUnstableException
    except Exception:
        return SimUnstable


def mesh_from_cells(cells):
    """The ``Cells`` attributes the loop consumes -> mesh dict (cells.py, SURVEY §2 ★data)."""
    m2c = np.asarray(cells.mem_to_cells)
    ptr = getattr(cells, "cell_mem_ptr", None)
    if ptr is None:
        C = len(cells.cell_vol)
        counts = np.bincount(m2c, minlength=C)
        ptr = np.concatenate(([0], np.cumsum(counts)))
        if not np.array_equal(np.repeat(np.arange(C), counts), m2c):
            raise BetseB200Error("membranes of a cell are not contiguous (cells.py:1095-1146 violated)")
    mesh = {"mem_to_cells": m2c, "cell_mem_ptr": np.asarray(ptr)}
    for f in ("nn_i", "bflags_mems", "map_mem2ecm", "mem_sa", "R_rads", "cell_vol", "cell_sa", "diviterm",
              "num_mems", "memSa_per_envSquare", "gj_default_weights"):
        mesh[f] = np.asarray(getattr(cells, f))
    if getattr(cells, "map_cell2ecm", None) is not None:
        mesh["map_cell2ecm"] = np.asarray(cells.map_cell2ecm)    # cell-zone rate laws reading env concentrations
    if getattr(cells, "mem_vol", None) is not None:
        mesh["mem_vol"] = np.asarray(cells.mem_vol)          # transporters' membrane-value nudge (networks.py:3020-3022)
    mv = getattr(cells, "mem_vects_flat", None)
    if mv is not None:
        mesh["mem_nx"], mesh["mem_ny"] = np.asarray(mv[:, 2]), np.asarray(mv[:, 3])
    else:
        mesh["mem_nx"], mesh["mem_ny"] = np.asarray(cells.mem_nx), np.asarray(cells.mem_ny)
    mesh["delta"] = np.asarray(float(cells.delta))
    mesh["gj_len"] = np.asarray(float(cells.gj_len))
    mesh["ecm_vol"] = np.asarray(float(cells.ecm_vol))
    X = getattr(cells, "X", None)
    mesh["grid_shape"] = np.asarray(X.shape if X is not None else cells.grid_shape)
    return mesh


def params_from_p(p):
    out = {f: getattr(p, f) for f in _P_FIELDS}
    ions = getattr(p, "ions", None)
    if ions is None:
        ions = [k for k, v in p.ions_dict.items() if v == 1]
    out["ions"] = np.array([str(x) for x in ions])
    return out


def state_from_sim(sim):
    out = {}
    for f in _STATE_FIELDS:
        if hasattr(sim, f):
            v = getattr(sim, f)
            if v is not None and not isinstance(v, dict):
                out[f] = np.asarray(v)
    bv = getattr(sim, "bound_V", None)
    if isinstance(bv, dict):
        out["bound_V"] = np.array([bv["T"], bv["B"], bv["L"], bv["R"]], dtype=float)
    elif bv is not None:
        out["bound_V"] = np.asarray(bv, dtype=float)
    return out


def _handlers(sim, p):
    """(handler index, MasterOfNetworks) of the enabled network handlers, in the order the loop runs them
    (sim.py:1290-1318): 0 = general network (sim.molecules.core), 1 = gene regulatory network (sim.grn.core)."""
    out = []
    if bool(getattr(p, "molecules_enabled", False)) and getattr(getattr(sim, "molecules", None), "core", None) is not None:
        out.append((0, sim.molecules.core))
    if bool(getattr(p, "grn_enabled", False)) and getattr(getattr(sim, "grn", None), "core", None) is not None:
        out.append((1, sim.grn.core))
    return out


def channels_from_sim(sim, p=None):
    """Channel specs of the network handlers (``core.channels``, built by ``Channel.init_channel``,
    networks.py:6550-6629) with their current gate states, in application order."""
    handlers = _handlers(sim, p) if p is not None else \
        [(0, core) for core in [getattr(getattr(sim, "molecules", None), "core", None)] if core is not None]
    out = []
    for h, core in handlers:
        for name, chan in (getattr(core, "channels", None) or {}).items():
            cc = chan.channel_core
            out.append({"name": name, "model": type(cc).__name__, "ion": cc.ions[0], "ions": list(cc.ions),
                        "maxDm": float(chan.maxDm), "rel_perm": float(cc.rel_perm[0]),
                        "rel_perms": [float(x) for x in cc.rel_perm], "init_active": bool(chan.init_active),
                        "targets": None if cc.targets is None else np.asarray(cc.targets),
                        "m": np.asarray(cc.m, dtype=float), "h": np.asarray(cc.h, dtype=float), "_obj": chan,
                        "handler": h, "mod_prog": -1})
    return out


def _bath_event_ions(sim, p):
    """Ion indices whose bath concentration a global event rewrites every step when the tissue has no extracellular
    spaces (tishandler.py:745-777: ``sim.cc_env[ion][:] = ...`` under K_env / Cl_env / Na_env)."""
    go = getattr(p, "global_options", None) or {}
    out = []
    if go.get("K_env", 0) != 0:
        out.append(int(sim.iK))
    if go.get("Cl_env", 0) != 0 and getattr(p, "ions_dict", {}).get("Cl", 0) == 1:
        out.append(int(sim.iCl))
    if go.get("Na_env", 0) != 0:
        out.append(int(sim.iNa))
    return out


def check_supported(sim, p):
    """Refuse loudly instead of silently computing a different model."""
    from . import network as netlib
    bad = []
    for h, core in _handlers(sim, p):
        bad += ["%s: %s" % ("general network" if h == 0 else "gene regulatory network", why)
                for why in netlib.unsupported_reasons(core, p)]
    for flag, what in (("deformation", "deformation"), ("deform_osmo", "osmotic pressure"),
                       ("fluid_flow", "fluid flow"), ("Ca_dyn", "ER calcium dynamics")):
        if bool(getattr(p, flag, False)):
            bad.append(what)
    if bad:
        raise BetseB200Error("betse_b200 does not implement: " + "; ".join(bad) +
                             " — run this configuration with the reference solver")


def engine_from_sim(sim, cells, p, device=0, phase_init=False):
    from . import network as netlib
    from . import ratelaw
    check_supported(sim, p)
    eng = TissueEngine(mesh_from_cells(cells), params_from_p(p), state_from_sim(sim), device=device)
    eng.chan_specs = []
    eng.net_cores = {}
    handlers = _handlers(sim, p)
    eng.all_cores = [core for _, core in handlers]
    if handlers:
        specs = channels_from_sim(sim, p)
        for h, core in handlers:
            if len(getattr(core, "molecules", None) or {}) == 0:
                continue
            desc = netlib.describe_core(core, sim, p, cells, record_static=False)
            comp = netlib.compile_network(desc, eng.Co, eng.M, ratelaw.live_resolver(core, sim, p, cells))
            eng.set_network(comp, handler=h)
            eng.net_cores[h] = core
            if comp.get("events"):
                eng.net_events = getattr(eng, "net_events", {})
                eng.net_events[h] = (desc, [None, None])        # description + the values last pushed
            for c in specs:
                if c["handler"] == h:
                    c["mod_prog"] = comp["mod_index"][comp["chan_names"].index(c["name"])]
        for c in specs:
            if c["handler"] not in eng.net_cores and c["_obj"].alpha_eval_string.replace(" ", "") not in (
                    "((np.ones(sim.mdl))*(np.ones(sim.mdl)))",):
                raise BetseB200Error("channel %r is modulated but its network has no substances" % c["name"])
        eng.set_channels(specs, phase_init=phase_init, affect_charge=bool(getattr(p, "substances_affect_charge", False)))
        eng.chan_specs = [c for c in specs if not (phase_init and not c["init_active"])]
    return eng


def _sample_fields(is_ecm, diag, noecm_field=False):
    """The per-sample download: what write2storage reads."""
    noecm_field = diag and not is_ecm and noecm_field
    fields = _W2S_STATE + (_W2S_ENV if is_ecm else ["cc_env"]) + (_W2S_DIAG if diag else [])
    if noecm_field:
        fields += ["v_env"]                    # venv_time: the local field potential (ion_current.py:158)
    if diag and (is_ecm or noecm_field):
        fields += ["J_env_x", "J_env_y"]       # I_tot_x_time / I_tot_y_time (sim.py:1866-1867)
    return fields


def _prefetch_staging(cells, p, device):
    """Start pinning the sampled-step staging while the engine is being built (ECM tissues; anything unexpected:
    no prefetch, the engine allocates on first use)."""
    try:
        if not bool(p.is_ecm):
            return None
        I = len(params_from_p(p)["ions"])
        X = getattr(cells, "X", None)
        E = int(np.prod(X.shape if X is not None else cells.grid_shape))
        dims = {"I": I, "C": len(cells.cell_vol), "M": len(cells.mem_sa), "E": E}
        shapes = {f: tuple(dims[k] for k in _DOWN_SHAPES[f]) for f in _sample_fields(True, True)}
        return PinnedPrefetch(device, shapes)
    except Exception:
        return None


def _copy_back(sim, eng, diag, sample_only=False):
    """Device state -> Simulator attributes.  ``sample_only``: just what write2storage reads, through the
    engine's page-locked staging (views that the next sample overwrites — write2storage copies what it
    keeps; vm_ave, which it appends as is (sim.py:1877), gets its own array)."""
    if sample_only:
        fields = _sample_fields(eng.is_ecm, diag, getattr(eng, "noecm_field", False))
    else:
        fields = list(_SAMPLED_STATE)
        if getattr(eng, "net_cores", None) or getattr(eng, "chan_specs", None):
            # what the handlers published last (networks.py:2971-2977): the NEXT phase's update_V reads them on the host
            fields += ["extra_rho_cells", "extra_J_mem"] + (["extra_rho_env"] if eng.is_ecm else [])
        if eng.is_ecm:
            fields += _SAMPLED_ENV
        if diag:
            fields += _SAMPLED_DIAG + ["E_gj_x", "E_gj_y"] + (_SAMPLED_DIAG_ENV if eng.is_ecm else [])
            if not eng.is_ecm and getattr(eng, "noecm_field", False):
                fields += ["v_env", "E_env_x", "E_env_y", "J_env_x", "J_env_y", "B_field", "Jtx", "Jty"]
    got = eng.download(fields, pinned=sample_only)
    shp = (eng.ny, eng.nx)
    for f, a in got.items():
        if f in ("E_env_x", "E_env_y", "J_env_x", "J_env_y", "Jtx", "Jty"):
            a = a.reshape(shp)                     # the reference keeps these 2-D (sim.py:572-573; sim_toolbox.py:1266-1290)
        if sample_only and f == "vm_ave":
            a = a.copy()
        elif sample_only:
            eng.__dict__.setdefault("_lent", {})[f] = a          # a view of engine-owned staging, see _detach
        setattr(sim, f, a)
    if not sample_only and callable(getattr(eng, "tj_modulator", None)):
        tjm = eng.tj_modulator()                    # tight-junction modulators rewrite it inside the loop (networks.py:3301-3317)
        if tjm is not None:
            sim.TJ_modulator = tjm.reshape(np.shape(sim.TJ_modulator)) if hasattr(sim, "TJ_modulator") else tjm
    # channel objects keep their gate state / open probability / flux (read by the exporters and by
    # the next phase through the pickled Simulator)
    for k, c in enumerate(getattr(eng, "chan_specs", [])):
        cc = c["_obj"].channel_core
        stt = eng.channel_state(k)
        tg = slice(None) if cc.targets is None else np.asarray(cc.targets)
        cc.m, cc.h, cc.P, cc.chan_flux, cc.DChan = stt["m"][tg], stt["h"][tg], stt["P"], stt["flux"], stt["DChan"]
    # network substances (read by MasterOfNetworks.write_data, networks.py:4210-4260, and the exporters)
    for h, core in getattr(eng, "net_cores", {}).items():
        m2c = eng.mem_to_cells.astype(np.int64)
        c, rates = eng.network_state(h, rates=True)
        env_on = eng.networks[h].get("env_on")
        cenv = eng.network_env_state(h) if env_on is not None and np.any(env_on) else None
        intra = eng.networks[h].get("intra_on")
        cmem = eng.network_mem_state(h) if intra is not None and np.any(intra) else None
        for k, name in enumerate(eng.networks[h]["species"]):
            mol = core.molecules[name]
            mol.c_cells = c[k].copy()
            mol.cc_at_mem = cmem[k].copy() if cmem is not None and intra[k] else c[k][m2c]
            if cenv is not None and env_on[k]:
                mol.c_env = cenv[k].copy()
        for desc_h, last in ([getattr(eng, "net_events", {}).get(h)] if h in getattr(eng, "net_events", {}) else []):
            if last[0] is not None:         # Molecule.update_boundary leaves the ramped boundary value on the object
                for ev in desc_h.get("events", []):
                    if ev["bounds"] is not None:
                        core.molecules[desc_h["species"][ev["species"]]].c_bound = float(last[0][ev["species"]])
        nk = len(eng.networks[h]["species"])
        core.reaction_rates = rates[nk:].copy()
    # MasterOfNetworks.energy_charge (networks.py:3996-4012), the tail of run_loop; write_data appends it
    # (networks.py:4256) whether or not the network has substances
    for core in getattr(eng, "all_cores", []):
        mols = getattr(core, "molecules", None) or {}
        if "AMP" in mols:
            cc = core.cell_concs
            core.chi = (cc["ATP"] + 0.5 * cc["ADP"]) / (cc["ATP"] + cc["ADP"] + cc["AMP"])
        else:
            core.chi = np.zeros(len(eng.mem_to_cells) and eng.Co)
    return 0


PIN_MIN_SAMPLES = 8
_closing = []        # engines being torn down on helper threads


def _join_closing():
    while _closing:
        _closing.pop().join()


def _close_async(eng):
    """Release an engine the loop owns without making the caller wait: freeing several GB of device memory and
    unpinning the staging takes 0.05-1 s (measured), none of which the Simulator needs — its arrays are complete when the
    loop returns.  The next engine creation and interpreter exit wait for it."""
    import atexit
    import threading
    if os.environ.get("BETSE_ASYNC_CLOSE", "1") == "0":
        eng.close()
        return
    if not getattr(_close_async, "registered", False):
        atexit.register(_join_closing)
        _close_async.registered = True
    th = threading.Thread(target=eng.close)
    th.start()
    _closing.append(th)


def _detach(sim, eng, own_engine=False):
    """Simulator attributes still aliasing the engine's page-locked staging: an engine that dies now leaves them the
    memory (engine._pinned_array frees it with the last view); one that lives on will overwrite its staging, so they get
    their own copy."""
    for f, a in ({} if own_engine else getattr(eng, "_lent", {})).items():
        if getattr(sim, f, None) is a:
            lib = getattr(eng, "lib", None)
            if lib is not None and a.flags.c_contiguous:
                own = np.empty_like(a)
                lib.betse_host_copy(own.ctypes.data, a.ctypes.data, a.nbytes)
            else:
                own = np.array(a, copy=True)
            setattr(sim, f, own)
    eng.__dict__["_lent"] = {}


def _dist_world():
    """torch.distributed, if this process is one rank of an initialised multi-rank group (torchrun), else None."""
    if os.environ.get("BETSE_STRIPS", "1") == "0":
        return None
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            return dist
    except Exception:
        pass
    return None


def _run_strips(sim, phase, time_steps, time_steps_sampled, anim_cells, dist, stats):
    """The loop over N GPUs: every rank of the process group runs the reference's host code on the WHOLE
    Simulator (events, sampling, storage — a few scalars per step) and steps ONE strip of the tissue on its GPU
    (strips.DistributedStrips: halo exchange inside the kernels, no collective on the data path).  At a sampled step every
    rank downloads its strip and the strips are all-gathered (plumbing: torch.distributed objects), so that every rank's
    Simulator holds the whole sampled state and `write2storage` stores the same series everywhere.  The ion path only:
    channels, networks, a boundary-voltage potential, polarizability, dynamic noise and the Helmholtz-Hodge diagnostics are
    refused on a decomposed tissue (`J_env_x/y`, `B_field`, `Jtx/y` keep their loop-entry values)."""
    from .partition import gather, ownership
    from .strips import DistributedStrips
    p, cells = phase.p, phase.cells
    check_supported(sim, p)
    if _handlers(sim, p):
        raise BetseB200Error("betse_b200: networks / channels on a domain-decomposed tissue are not implemented")
    if not bool(p.is_ecm):
        raise BetseB200Error("betse_b200: domain decomposition needs extracellular spaces (run one replica per GPU instead)")
    if anim_cells is not None:
        raise BetseB200Error("betse_b200: mid-simulation animations are not available on a domain-decomposed tissue")
    kind = getattr(getattr(phase, "kind", None), "name", str(getattr(phase, "kind", "")))
    is_sim = kind.upper() == "SIM"
    fire = getattr(getattr(phase, "dyna", None), "fire_events", None) if is_sim else None
    if is_sim and getattr(p, "dynamic_noise", False) == 1:
        raise BetseB200Error("betse_b200: dynamic noise on a domain-decomposed tissue is not implemented")
    t0 = time.time()
    _join_closing()
    mesh, params, state = mesh_from_cells(cells), params_from_p(p), state_from_sim(sim)
    R, rank = dist.get_world_size(), dist.get_rank()
    ds = DistributedStrips(mesh, params, state, int(os.environ.get("LOCAL_RANK", rank)), dist)
    own = ownership(mesh, R)
    Unstable = _unstable_exception()
    sampled = set(time_steps_sampled)
    cache = {f: np.array(getattr(sim, f), copy=True) for f in _live_scheduled(p) if hasattr(sim, f)
             and getattr(sim, f) is not None} if fire else {}
    hh = ("J_env_x", "J_env_y", "B_field", "Jtx", "Jty")

    def copy_back(diag):
        fields = [f for f in (_sample_fields(True, diag) if diag else list(_SAMPLED_STATE) + _SAMPLED_ENV) if f not in hh]
        mine = ds.download_local(fields)
        allf = [None] * R
        dist.all_gather_object(allf, mine)
        got = gather(own, allf)
        shp = (int(mesh["grid_shape"][0]), int(mesh["grid_shape"][1]))
        for f, a in got.items():
            if f in ("E_env_x", "E_env_y"):
                a = a.reshape(shp)
            setattr(sim, f, a)
        return sum(x.nbytes for x in mine.values())
    n, n_total, d2h = 0, len(time_steps), 0
    try:
        while n < n_total:
            if fire is not None:
                fire(phase=phase, t=time_steps[n])
                for f in list(cache):
                    new = np.asarray(getattr(sim, f))
                    if new.shape != cache[f].shape or not np.array_equal(new, cache[f]):
                        ds.set_field(f, new)
                        cache[f] = np.array(new, copy=True)
                run = 1
            else:
                run = 1
                while n + run < n_total and time_steps[n + run - 1] not in sampled:
                    run += 1
            last_t = time_steps[n + run - 1]
            is_sampled = last_t in sampled
            status = ds.step(run, diag=is_sampled)
            n += run
            # an instability anywhere stops every rank (the status word is local to a strip)
            flags = [None] * R
            dist.all_gather_object(flags, int(status))
            status = 0
            for x in flags:
                status |= x
            if status & (capi.STATUS_NAN_VM | capi.STATUS_NAN_CONC):
                d2h += copy_back(False)
                raise Unstable("Your simulation has become unstable. Please try a smaller time step,"
                               "reduce gap junction radius, and/or reduce pump rate coefficients.")
            if is_sampled:
                d2h += copy_back(True)
                phase.callbacks.progressed_next()
                sim.write2storage(last_t, cells, p)
        if n_total and time_steps[n_total - 1] not in sampled:
            d2h += copy_back(False)
    finally:
        if stats is not None:
            stats.update({"h2d_bytes": ds.engine.h2d_bytes, "d2h_bytes": d2h, "wall_s": time.time() - t0, "steps": n,
                          "n_gpus": R})
        ds.close()


def run_sim_core_loop(sim, phase, time_steps, time_steps_sampled, anim_cells=None, *,
                      engine=None, device=0, stats=None):
    """Same signature and contract as Simulator._run_sim_core_loop (sim.py:1132-1138).  Under torchrun (an initialised
    torch.distributed group of N > 1 ranks) the tissue is stepped as N strips, one per GPU: see _run_strips."""
    p, cells = phase.p, phase.cells
    dist = _dist_world() if engine is None else None
    if dist is not None:
        ev_cut = getattr(getattr(phase, "dyna", None), "event_cut", None)
        kind0 = getattr(getattr(phase, "kind", None), "name", str(getattr(phase, "kind", "")))
        if kind0.upper() == "SIM" and ev_cut is not None and not ev_cut.is_fired and len(time_steps):
            phase.dyna.fire_events(phase=phase, t=time_steps[0])       # the cutting event, as below
        return _run_strips(sim, phase, time_steps, time_steps_sampled, anim_cells, dist, stats)
    own_engine = engine is None
    kind = getattr(getattr(phase, "kind", None), "name", str(getattr(phase, "kind", "")))
    is_sim = kind.upper() == "SIM"
    fire = getattr(getattr(phase, "dyna", None), "fire_events", None) if is_sim else None
    ev_cut = getattr(getattr(phase, "dyna", None), "event_cut", None)
    if fire is not None and engine is None and ev_cut is not None and not ev_cut.is_fired and len(time_steps):
        # The cutting event (tishandler.py:884-913) fires inside the FIRST fire_events call of the phase
        # (event_cut_time is hard-wired to 0, parameters.py:686-687) and re-indexes the mesh and every state array
        # on the host (TissueHandler._cut_cells, tishandler.py:921-1285: the reference's own code).  Nothing has run on
        # the device yet, so let it happen before the engine is built from the (then post-cut) Simulator; the loop's
        # own call for this step finds the event fired and only re-evaluates the scheduled scalars for the same t.
        fire(phase=phase, t=time_steps[0])
        cells = phase.cells
    t0 = time.time()
    tm = {"engine": 0.0, "steps": 0.0, "samples": 0.0, "final": 0.0, "close": 0.0}
    sampled = set(time_steps_sampled)
    if engine is None:
        _join_closing()                  # the previous phase's engine has released its device memory
        # page-locked staging for the sampled-step downloads costs ~0.2 s per 435 MB to set up and saves ~40 ms per
        # sample: worth it from PIN_MIN_SAMPLES samples on (measured, profiles/r02g_e2e_fixed_cost.txt)
        pin = len(sampled) >= PIN_MIN_SAMPLES
        pf = _prefetch_staging(cells, p, device) if (
            anim_cells is None and pin and hasattr(TissueEngine, "adopt_pinned")
            and os.environ.get("BETSE_PIN_PREFETCH", "1") != "0") else None
        try:
            eng = engine_from_sim(sim, cells, p, device=device, phase_init=not is_sim)
        except BaseException:
            if pf is not None:
                pf.release()
            raise
        if pf is not None:
            eng.adopt_pinned(pf)
        eng.pin_staging = pin
    else:
        eng = engine
    tm["engine"] = time.time() - t0
    noisy = is_sim and getattr(p, "dynamic_noise", False) == 1 and "P" in eng.ions
    Unstable = _unstable_exception()
    h2d = d2h = 0
    h2d0, d2h0 = eng.h2d_bytes, eng.d2h_bytes
    cache = {f: np.array(getattr(sim, f), copy=True) for f in _live_scheduled(p) if hasattr(sim, f)
             and getattr(sim, f) is not None} if fire else {}
    bv_cache = dict(getattr(sim, "bound_V", {})) if fire else {}
    bath_events = _bath_event_ions(sim, p) if (fire and not eng.is_ecm) else []
    n_total = len(time_steps)
    n = 0
    try:
        while n < n_total:
            if fire is not None:
                # scheduled interventions are host-side scalar logic: keep the reference's own code
                fire(phase=phase, t=time_steps[n])
                for f in list(cache):
                    new = np.asarray(getattr(sim, f))
                    if new.shape != cache[f].shape or not np.array_equal(new, cache[f]):
                        eng.set_field(f, new)
                        cache[f] = np.array(new, copy=True)
                if bath_events:
                    # without extracellular spaces the global K_env / Cl_env / Na_env events overwrite the bath
                    # concentration itself on every step (tishandler.py:759-777), whatever the step before made of it
                    eng.set_bath({i: float(np.asarray(sim.cc_env[i]).reshape(-1)[0]) for i in bath_events})
                bv = getattr(sim, "bound_V", None)
                if isinstance(bv, dict) and bv != bv_cache:
                    # the external-voltage event (tissue/event/tisevevolt.py:76-88): Phi_b is re-solved on the device
                    eng.set_bound_V([bv["T"], bv["B"], bv["L"], bv["R"]])
                    bv_cache = dict(bv)
                run = 1
            elif getattr(eng, "net_events", None) or noisy:
                run = 1                     # substances with timed events / dynamic noise: host input for every step
            else:
                # no events: run up to and including the next sampled step in one call
                run = 1
                while n + run < n_total and time_steps[n + run - 1] not in sampled:
                    run += 1
            # Molecule.update_boundary / cell_clamp_method (networks.py:2929-2933): host-side scalar schedules of the step's time
            for h, (desc_h, last) in getattr(eng, "net_events", {}).items():
                cb, cl = netlib_event_values(desc_h, float(time_steps[n]))
                if last[0] is None or not np.array_equal(cb, last[0]) or not np.array_equal(cl, last[1], equal_nan=True):
                    eng.set_network_events(h, cb, cl)
                    last[0], last[1] = cb, cl
            if noisy:
                # dynamic noise (sim.py:1322-1339): the draw comes from NumPy's global stream, like the reference's — nothing
                # else in the loop draws from it, so the sequence is the one the reference would have used
                sim.protein_noise_flux = p.dynamic_noise_level * (np.random.random(eng.M) - 0.5)
                eng.set_noise_flux(sim.protein_noise_flux)
            last_t = time_steps[n + run - 1]
            is_sampled = last_t in sampled
            t1 = time.time()
            status = eng.step(run, diag=is_sampled)
            tm["steps"] += time.time() - t1
            n += run
            if status & capi.STATUS_NEG_NET:
                d2h += _copy_back(sim, eng, diag=False)
                raise Unstable("Network concentration in cells below zero! Your simulation has become unstable.")
            if status & (capi.STATUS_NAN_VM | capi.STATUS_NAN_CONC):
                d2h += _copy_back(sim, eng, diag=False)
                raise Unstable(
                    "Your simulation has become unstable. Please try a smaller time step,"
                    "reduce gap junction radius, and/or reduce pump rate coefficients.")
            if is_sampled:
                # the last step of the phase leaves complete, engine-independent arrays on the Simulator
                final = n >= n_total
                t1 = time.time()
                d2h += _copy_back(sim, eng, diag=True, sample_only=(anim_cells is None and not final))
                tm["final" if final else "samples"] += time.time() - t1
                phase.callbacks.progressed_next()
                sim.write2storage(last_t, cells, p)
                if anim_cells is not None:
                    anim_cells.plot_frame(time_step=-1)
        # leave the Simulator holding the final state, like the reference does
        if n_total and time_steps[n_total - 1] not in sampled:
            t1 = time.time()
            d2h += _copy_back(sim, eng, diag=False)
            tm["final"] += time.time() - t1
    finally:
        h2d = eng.h2d_bytes - (0 if own_engine else h2d0)
        d2h = eng.d2h_bytes - (0 if own_engine else d2h0)
        t1 = time.time()
        _detach(sim, eng, own_engine)
        if own_engine:
            _close_async(eng)
        tm["close"] = time.time() - t1
        if stats is not None:
            stats.update({"h2d_bytes": h2d, "d2h_bytes": d2h, "wall_s": time.time() - t0, "steps": n,
                          "seconds": {k: round(v, 4) for k, v in tm.items()}})


def run_fast_sim_core_loop(sim, phase, time_steps, time_steps_sampled, anim_cells=None, *, engine=None, device=0, stats=None):
    """Same signature and contract as Simulator._run_fast_sim_core_loop (sim.py:1454-1640), the equivalent-circuit
    solver selected by ``solver options: type: fast``: the loop body runs on the device (csrc/fast.cu), scheduled events
    stay the reference's own host code, and the sampled steps append to the Simulator's time series exactly what the
    reference's loop appends (sim.py:1597-1628).  Networks under the fast solver: voltage-gated channels
    (run_fast_loop_channels, networks.py:3217-3280) of handlers WITHOUT substances, transporters or modulators; the rest
    (their run_loop moves concentrations the equivalent circuit never reads back) is refused."""
    p, cells = phase.p, phase.cells
    kind = getattr(getattr(phase, "kind", None), "name", str(getattr(phase, "kind", "")))
    handlers = _handlers(sim, p)
    for h, core in handlers:
        if len(getattr(core, "molecules", None) or {}) or len(getattr(core, "transporters", None) or {}) or \
                len(getattr(core, "modulators", None) or {}) or len(getattr(core, "reactions", None) or {}):
            raise BetseB200Error("betse_b200: network substances / transporters / modulators under the fast solver "
                                 "(networks.py:2805-2982 inside sim.py:1505-1545) are not implemented")
    specs = channels_from_sim(sim, p) if handlers else []
    for c in specs:
        if c["_obj"].alpha_eval_string.replace(" ", "") not in ("((np.ones(sim.mdl))*(np.ones(sim.mdl)))",):
            raise BetseB200Error("betse_b200: channel %r is modulated: not implemented under the fast solver" % c["name"])
    fire = getattr(getattr(phase, "dyna", None), "fire_events", None) if kind.upper() == "SIM" else None
    own = engine is None
    t0 = time.time()
    if own:
        _join_closing()
        eng = TissueEngine(mesh_from_cells(cells), params_from_p(p), state_from_sim(sim), device=device)
    else:
        eng = engine
    eng.chan_specs = []
    if specs:
        phase_init = kind.upper() == "INIT"
        eng.set_channels(specs, phase_init=phase_init, affect_charge=False)
        eng.chan_specs = [c for c in specs if not (phase_init and not c["init_active"])]
        I = len(sim.zs)
        cbar, rev = np.zeros(I), np.zeros(I)
        for ion in sim.rev_E_dic:                       # Simulator.fast_sim_init, sim.py:1393-1452
            cbar[sim.get_ion(ion)] = float(np.mean(sim.cbar_dic[ion]))
            rev[sim.get_ion(ion)] = float(np.mean(sim.rev_E_dic[ion]))
        eng.fast_set_channels(cbar, rev, float(getattr(sim, "geo_conv", 1.0)))
    eng.fast_setup({f: getattr(sim, f, None) for f in ("vm_ave", "gjopen", "G_Leak", "E_Leak", "G_gj", "sigma_cell", "extra_J_mem")})
    Unstable = _unstable_exception()
    sampled = set(time_steps_sampled)
    gjb = np.array(getattr(sim, "gj_block", 1.0), dtype=float, copy=True)
    n, n_total = 0, len(time_steps)
    m2c = np.asarray(cells.mem_to_cells)

    def copy_back(diag):
        got = eng.fast_download(list(TissueEngine.FAST_FIELDS) if diag else ["vm_ave", "gjopen", "vgj"])
        for f, a in got.items():
            setattr(sim, f, a)
        if diag:
            sim.Jn = got["Jn"]
        # channel objects keep their gate state / open probability / flux (networks.py:3256-3280; read by the exporters)
        for k, c in enumerate(getattr(eng, "chan_specs", [])):
            cc = c["_obj"].channel_core
            stt = eng.channel_state(k)
            tg = slice(None) if cc.targets is None else np.asarray(cc.targets)
            cc.m, cc.h, cc.P, cc.chan_flux = stt["m"][tg], stt["h"][tg], stt["P"], stt["flux"]
    try:
        while n < n_total:
            if fire is not None:
                fire(phase=phase, t=time_steps[n])
                new = np.asarray(getattr(sim, "gj_block", 1.0), dtype=float)
                if new.shape != gjb.shape or not np.array_equal(new, gjb):
                    eng.set_field("gj_block", new)                 # the only scheduled quantity this solver reads
                    gjb = np.array(new, copy=True)
                run = 1
            else:
                run = 1
                while n + run < n_total and time_steps[n + run - 1] not in sampled:
                    run += 1
            last_t = time_steps[n + run - 1]
            is_sampled = last_t in sampled
            status = eng.fast_step(run, diag=is_sampled)
            n += run
            if status & capi.STATUS_NAN_VM:
                copy_back(False)
                raise Unstable("Your simulation has become unstable. Please try a smaller time step,"
                               "reduce gap junction radius, and/or reduce pump rate coefficients.")
            if is_sampled:
                copy_back(True)
                phase.callbacks.progressed_next()
                # sim.py:1603-1628
                sim.vm_time.append(sim.vm * 1)
                sim.dd_time.append(np.copy(sim.Dm_cells))
                sim.I_cell_x_time.append(sim.J_cell_x * 1)
                sim.I_cell_y_time.append(sim.J_cell_y * 1)
                sim.efield_gj_x_time.append(sim.E_cell_x[m2c] * 1)
                sim.efield_gj_y_time.append(sim.E_cell_y[m2c] * 1)
                sim.gjopen_time.append(sim.gjopen * 1)
                sim.time.append(last_t * 1)
                for _, core in handlers:                 # sim.py:1616-1622; chi: energy_charge at the tail of run_loop
                    core.chi = np.zeros(eng.Co)
                    core.write_data(sim, cells, p)
                    core.report(sim, p)
                sim.vm_ave_time.append(sim.vm_ave * 1)
                if anim_cells is not None:
                    anim_cells.plot_frame(time_step=-1)
        if n_total and time_steps[n_total - 1] not in sampled:
            # the reference forms the currents and fields on every step: leave the Simulator the last step's
            eng.fast_step(0, diag=True)
            copy_back(True)
        for _, core in handlers:
            core.chi = np.zeros(eng.Co)
    finally:
        if own:
            _close_async(eng)
        if stats is not None:
            stats.update({"wall_s": time.time() - t0, "steps": n, "h2d_bytes": eng.h2d_bytes, "d2h_bytes": eng.d2h_bytes})


_installed = None
_installed_fast = None


def install(device=0):
    """Rebind the reference's solvers to the B200 loops (sim.py:1064-1075 seam): the full solver and the fast one."""
    global _installed, _installed_fast
    from betse.science.sim import Simulator
    if _installed is None:
        _installed = Simulator._run_sim_core_loop
        _installed_fast = Simulator._run_fast_sim_core_loop

    def _loop(self, phase, time_steps, time_steps_sampled, anim_cells):
        return run_sim_core_loop(self, phase, time_steps, time_steps_sampled, anim_cells, device=device)

    def _fast(self, phase, time_steps, time_steps_sampled, anim_cells):
        return run_fast_sim_core_loop(self, phase, time_steps, time_steps_sampled, anim_cells, device=device)
    Simulator._run_sim_core_loop = _loop
    Simulator._run_fast_sim_core_loop = _fast


def uninstall():
    global _installed, _installed_fast
    if _installed is not None:
        from betse.science.sim import Simulator
        Simulator._run_sim_core_loop = _installed
        Simulator._run_fast_sim_core_loop = _installed_fast
        _installed = _installed_fast = None
