#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the REAL reference (/root/reference) through
oracle/refrun.py.  Run in the build container only:

    python tests/golden/make_golden.py [scenario ...]

Each fixture holds the mesh, parameters, loop-entry state and the reference's own outputs
after N timesteps of Simulator._run_sim_core_loop (betse/science/sim.py:1132-1390) for the
INIT phase (from the uniform initial state: vm == 0, vgj == 0 exactly — the degenerate
edge case) and the SIM phase (from the relaxed state after the full INIT run).
Seed: np.random.seed(12345) before `seed`.  Versions are recorded in each file.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import refrun  # noqa: E402

NO_NET = {"general network": {"implement network": False},
          "cutting event": {"event happens": False}}


def _m(*ds):
    out = {}

    def merge(d, m):
        for k, v in m.items():
            if isinstance(v, dict):
                merge(d.setdefault(k, {}), v)
            else:
                d[k] = v
    for d in ds:
        merge(out, d)
    return out


SMALL = {"world options": {"world size": 100e-6}, "general options": {"comp grid size": 20},
         "init time settings": {"total time": 1.0, "sampling rate": 0.1}}

SCENARIOS = {
    # the shipped `betse try` world (C=228, M=1309, 24x25 grid), ion path only
    "basic_ecm": dict(mods=_m(NO_NET), snaps={"init": [1, 2, 5, 20], "sim": [1, 2, 5, 20]}),
    "mammal_ecm": dict(mods=_m(NO_NET, SMALL, {"general options": {"ion profile": "mammal"}}),
                       snaps={"init": [1, 2, 5], "sim": [1, 2, 5, 20]}),
    # BASELINE configs[0]: Na/K/Cl/Ca/P(+M), no ECM
    "mammal_noecm": dict(mods=_m(NO_NET, SMALL, {"general options": {
        "ion profile": "mammal", "simulate extracellular spaces": False}}),
        snaps={"init": [1, 2, 5], "sim": [1, 2, 5, 20]}),
    "basic_noecm": dict(mods=_m(NO_NET, SMALL, {"general options": {
        "simulate extracellular spaces": False}}),
        snaps={"init": [1, 2], "sim": [1, 5]}),
    # non-default numerics switches: env smoothing, fast ecm exchange, static GJ, K noise
    "mammal_ecm_opts": dict(mods=_m(NO_NET, SMALL, {
        "general options": {"ion profile": "mammal"},
        "internal parameters": {"sharpness env": 0.9, "fast update ecm": True},
        "variable settings": {"gap junctions": {"voltage sensitive gj": False},
                              "noise": {"static noise level": 0.5},
                              "tight junction scaling": 0.5, "adherens junction scaling": 0.8}}),
        snaps={"init": [1, 2], "sim": [1, 5]}),
    "mammal_ecm_polar": dict(mods=_m(NO_NET, SMALL, {
        "general options": {"ion profile": "mammal"},
        "internal parameters": {"cell polarizability": 1.0e-4}}),
        snaps={"init": [1, 2], "sim": [1, 5]}),
}


def chan_extra(sim, phase):
    """Channel library state (networks.py:6550-6629; channels/*.py): model, targets, gates."""
    out = {}
    core = getattr(getattr(sim, "molecules", None), "core", None)
    if core is None or not getattr(core, "channels", None):
        return out
    names = list(core.channels)
    out["chan.names"] = np.array(names)
    for k, n in enumerate(names):
        c = core.channels[n]
        cc = c.channel_core
        out["chan%d.model" % k] = np.array(type(cc).__name__)
        out["chan%d.ion" % k] = np.array(cc.ions[0])
        out["chan%d.rel_perm" % k] = np.asarray(float(cc.rel_perm[0]))
        out["chan%d.maxDm" % k] = np.asarray(float(c.maxDm))
        out["chan%d.init_active" % k] = np.asarray(int(bool(c.init_active)))
        out["chan%d.targets" % k] = np.asarray(cc.targets, dtype=np.int64)
        out["chan%d.m" % k] = np.asarray(cc.m, dtype=float) * np.ones(len(cc.targets))
        out["chan%d.h" % k] = np.asarray(getattr(cc, "h", 1.0), dtype=float) * np.ones(len(cc.targets))   # Morris-Lecar: no h gate
        if hasattr(cc, "P"):
            out["chan%d.P" % k] = np.asarray(cc.P, dtype=float)
        if getattr(cc, "chan_flux", None) is not None:
            out["chan%d.flux" % k] = np.asarray(cc.chan_flux, dtype=float)
    if getattr(sim, "rev_E_dic", None):
        # constants of Simulator.fast_sim_init (sim.py:1393-1452) the fast solver's channels read, in ion-index order
        I = len(sim.zs)
        rev, cbar = np.zeros(I), np.zeros(I)
        for ion in sim.rev_E_dic:
            rev[sim.get_ion(ion)] = float(np.mean(sim.rev_E_dic[ion]))
            cbar[sim.get_ion(ion)] = float(np.mean(sim.cbar_dic[ion]))
        out["fast.rev_E"], out["fast.cbar"], out["fast.geo_conv"] = rev, cbar, np.asarray(float(sim.geo_conv))
    return out


CHANNELS = [
    {"name": "Nav", "channel class": "Na", "channel type": "Nav1p3", "max Dm": 2.0e-14, "apply to": "all", "init active": False},
    {"name": "Kv", "channel class": "K", "channel type": "Kv1p5", "max Dm": 1.0e-15, "apply to": "all", "init active": False},
    {"name": "K_Leak", "channel class": "K", "channel type": "KLeak", "max Dm": 0.6e-17, "apply to": "all", "init active": True},
    {"name": "Cav", "channel class": "Ca", "channel type": "Cav1p2", "max Dm": 1.0e-15, "apply to": "all", "init active": False},
]
# BASELINE configs[2] in small: voltage-gated Na/K/Ca channels + pumps, full ion profile, ECM
SCENARIOS["mammal_ecm_chan"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": [], "channels": CHANNELS}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=chan_extra)


# multi-ion families: a hyperpolarisation-activated HCN2 (vg_funny.py: Na/K/Ca through one gate) on the whole tissue
# and a non-selective cation leak (cation.py) on one tissue profile, next to a voltage-gated K channel
CHANNELS_MULTI = [
    {"name": "Funny", "channel class": "Fun", "channel type": "HCN2", "max Dm": 5.0e-16, "apply to": "all", "init active": True},
    {"name": "Kv", "channel class": "K", "channel type": "Kv1p5", "max Dm": 1.0e-15, "apply to": "all", "init active": False},
    {"name": "Leaky", "channel class": "Cat", "channel type": "CatLeak2", "max Dm": 2.0e-18, "apply to": ["Spot"], "init active": True},
]
SCENARIOS["mammal_ecm_chan_multi"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": [], "channels": CHANNELS_MULTI}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=chan_extra)


# Morris-Lecar family (vg_morrislecar.py, `channel class: ML`, networks.py:6611-6613): a kinetic K channel (update_ml), an
# instantaneous Na gate (m = mInf) and the two-ion HCN2_ML, next to a Hodgkin-Huxley Ca channel
CHANNELS_ML = [
    {"name": "KvML", "channel class": "ML", "channel type": "Kv1p5_ML", "max Dm": 1.0e-17, "apply to": "all", "init active": True},
    {"name": "NavML", "channel class": "ML", "channel type": "Nav_ML", "max Dm": 2.0e-17, "apply to": "all", "init active": False},
    {"name": "FunnyML", "channel class": "ML", "channel type": "HCN2_ML", "max Dm": 5.0e-18, "apply to": ["Spot"], "init active": True},
    {"name": "Cav", "channel class": "Ca", "channel type": "Cav1p2", "max Dm": 1.0e-15, "apply to": "all", "init active": False},
]
SCENARIOS["mammal_ecm_chan_ml"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": [], "channels": CHANNELS_ML}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=chan_extra)


# the FAST (equivalent-circuit) solver, Simulator._run_fast_sim_core_loop (sim.py:1454-1640): no networks — gap-junction
# coupled leak circuit with voltage-sensitive gap junctions; with and without extracellular spaces (fast_sim_init reads
# the env concentrations either way)
def _leaky_spot(p):
    """A K-leaky tissue profile: its cells rest at another potential (vm_GHK -> E_Leak), so that current flows through
    the gap junctions and their voltage gating moves."""
    for t in p.tissue_profiles:
        if t.name == "Spot":
            t.Dm_K = 3.0e-17


SCENARIOS["fast_basic"] = dict(
    mods=_m(NO_NET, SMALL, {"solver options": {"type": "fast"}}), tweak_p=_leaky_spot,
    snaps={"init": [1, 2, 5, 20], "sim": [1, 2, 5, 20]}, method="_run_fast_sim_core_loop")
SCENARIOS["fast_mammal_noecm"] = dict(
    mods=_m(NO_NET, SMALL, {"solver options": {"type": "fast"}, "general options": {"ion profile": "mammal", "simulate extracellular spaces": False}}),
    tweak_p=_leaky_spot, snaps={"init": [1, 2, 5, 20], "sim": [1, 2, 5, 20]}, method="_run_fast_sim_core_loop")


# the fast solver with voltage-gated channels (run_fast_loop_channels, networks.py:3217-3280): conductances from the open
# probabilities, currents against the fixed reversal potentials of fast_sim_init, joined into extra_J_mem
SCENARIOS["fast_chan"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "solver options": {"type": "fast"},
                    "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": [], "channels": CHANNELS}}),
    tweak_p=_leaky_spot, snaps={"init": [1, 2, 5, 20], "sim": [1, 2, 5, 20]}, method="_run_fast_sim_core_loop", extra=chan_extra)


def _substance(name, prod, acts=None, inh=None, Dgj=1e-15, gj_imp=True, cell=0.1, z=0, apply_to="all"):
    gd = {"production rate": prod, "decay rate": 1.0, "apply to": apply_to, "modulator function": "None"}
    if acts:
        gd.update({"activators": [a for a, _, _ in acts], "Km activators": [k for _, k, _ in acts],
                   "n activators": [n for _, _, n in acts]})
    if inh:
        gd.update({"inhibitors": [a for a, _, _ in inh], "Km inhibitors": [k for _, k, _ in inh],
                   "n inhibitors": [n for _, _, n in inh]})
    return {"name": name, "Dm": 0.0, "Do": 1.0e-10, "Dgj": Dgj, "z": z, "env conc": 0.0, "cell conc": cell,
            "scale factor": 1.0, "update intracellular": False, "use time dilation": False, "transmem": False,
            "initial asymmetry": "None", "TJ permeable": False, "GJ impermeable": gj_imp, "TJ factor": 1.0,
            "growth and decay": gd,
            "plotting": {"plot 2D": False, "animate": False, "autoscale colorbar": True, "max val": 2.0, "min val": 0.0}}


def net_extra(sim, phase):
    """General network state (MasterOfNetworks, networks.py): the description betse_b200.network
    compiles (strings, tables, static terms) + substance concentrations, rates and channel DChan."""
    from betse_b200 import network as netlib
    out = chan_extra(sim, phase)
    p = phase.p
    for h, holder in ((0, getattr(sim, "molecules", None) if p.molecules_enabled else None),
                      (1, getattr(sim, "grn", None) if p.grn_enabled else None)):     # sim.py:1290-1319: general network, then GRN
        core_h = getattr(holder, "core", None)
        if core_h is None or not len(getattr(core_h, "molecules", None) or {}):
            continue
        desc = netlib.describe_core(core_h, sim, p, phase.cells)
        out.update(netlib.flatten(desc, "net%d." % h))
        if getattr(core_h, "reaction_rates", None) is not None and len(core_h.reaction_rates):
            out["net%d.reaction_rates" % h] = np.asarray(core_h.reaction_rates, dtype=float)
    core = sim.molecules.core if p.molecules_enabled and getattr(sim, "molecules", None) is not None else None
    if core is None:
        return out
    for k, n in enumerate(core.channels):
        cc = core.channels[n].channel_core
        if getattr(cc, "DChan", None) is not None:
            out["chan%d.DChan" % k] = np.asarray(cc.DChan, dtype=float) * np.ones(sim.mdl)
    return out


# BASELINE configs[3] in small: a gene-regulatory style network (Hill activation / inhibition, the shape of
# extra_configs/grn_basic.yaml) + a gap-junction permeable substance grown in one tissue profile (the shipped
# default's 'X') + a cell-zone reaction, coupled to Vmem through substance-modulated K channels
_NET_BIO = [_substance("G1", 2.0, inh=[("G3", 0.01, 1)]), _substance("G2", 2.0, acts=[("G1", 1, 1)]),
            _substance("G3", 15.0, acts=[("G1", 1, 1), ("G2", 1, 2.5)]),
            _substance("X", 0.1, Dgj=1e-15, gj_imp=False, cell=0.05, apply_to=["Spot"])]
_NET_RX = [{"name": "make_X", "reaction zone": "cell", "reactants": ["G2"], "reactant multipliers": [1],
            "Km reactants": [1.0], "products": ["X"], "product multipliers": [1], "Km products": [0.1],
            "max rate": 5.0e-3, "standard free energy": "None"}]
_NET_CH = [dict(c) for c in CHANNELS]
_NET_CH[2] = dict(_NET_CH[2], **{"channel inhibitors": ["X"], "inhibitor Km": [0.05], "inhibitor n": [2.0],
                                 "inhibitor zone": ["cell"]})
_NET_CH[1] = dict(_NET_CH[1], **{"channel activators": ["G3"], "activator Km": [0.5], "activator n": [1.0],
                                 "activator zone": ["cell"]})
SCENARIOS["mammal_ecm_net"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _NET_BIO, "reactions": _NET_RX,
                                        "channels": _NET_CH}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# Substances that cross the membrane and diffuse / electro-migrate through the extracellular grid (Molecule.transport ->
# stb.molecule_mover, sim_toolbox.py:909-1153): S1 neutral, membrane- and gap-junction permeable, held back by tight
# junctions; S2 extracellular only (bath + boundary), passes tight junctions; S3 a cation with a membrane permeability;
# G1 regulated by S1 inside the cell.  'substances affect Vmem' off: charged substances then leave Vmem alone.
def _env_substance(name, Dm, z, env, cell, gj_imp, tj_perm, tj_factor=1.0, prod=0.0, inh=None):
    s = _substance(name, prod, inh=inh, Dgj=1e-15, gj_imp=gj_imp, cell=cell, z=z)
    s.update({"Dm": Dm, "env conc": env, "TJ permeable": tj_perm, "TJ factor": tj_factor})
    return s


_ENV_BIO = [_env_substance("S1", 2.0e-17, 0, 0.5, 0.1, False, False, tj_factor=0.5),
            _env_substance("S2", 0.0, 0, 0.2, 0.0, True, True),
            _env_substance("S3", 5.0e-18, 1, 0.3, 0.05, True, False),
            _substance("G1", 2.0, inh=[("S1", 0.2, 2)])]
SCENARIOS["mammal_ecm_net_env"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "internal parameters": {"substances affect Vmem": False},
                    "general network": {"implement network": True, "biomolecules": _ENV_BIO, "reactions": [],
                                        "channels": [dict(_NET_CH[2], **{"channel inhibitors": ["S1"]})]}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# the same with 'substances affect Vmem' on (the shipped default): the cation S3 and an anion S4 (scale factor 0.5) add to
# the charge of cells and env squares, their membrane / gap-junction fluxes to the membrane current (networks.py:2942-2977)
_ENVQ_BIO = [dict(b) for b in _ENV_BIO[:3]] + [dict(_env_substance("S4", 1.0e-17, -1, 0.1, 0.4, False, True), **{"scale factor": 0.5}),
                                               _substance("G1", 2.0, inh=[("S1", 0.2, 2)])]
SCENARIOS["mammal_ecm_net_envq"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _ENVQ_BIO, "reactions": [],
                                        "channels": [dict(_NET_CH[2], **{"channel inhibitors": ["S1"]})]}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# sim modulators (run_loop_modulators, networks.py:3282-3325): G3 closes gap junctions, X drives the Na/K-ATPase
_MODS = [{"name": "gj_mod", "target": "GJ", "max effect": 1.0, "inhibitors": ["G3"], "inhibitor Km": [2.0], "inhibitor n": [2.0],
          "inhibitor zone": ["cell"]},
         {"name": "pump_mod", "target": "Na/K-ATPase", "max effect": 1.5, "activators": ["X"], "activator Km": [0.05],
          "activator n": [1.0], "activator zone": ["cell"]}]
SCENARIOS["mammal_ecm_net_mod"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _NET_BIO, "reactions": _NET_RX,
                                        "channels": _NET_CH, "modulators": _MODS}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# tight-junction modulators (run_loop_modulators, target 'TJ', networks.py:3301-3317): an extracellular-zone rate law
# (get_influencers with reaction_zone 'env', networks.py:5270-5281) rewrites sim.TJ_modulator on the tight-junction squares
# every step — S1 in the bath opens the barrier for all ions, S3 closes it for Na only
_TJ_MODS = [{"name": "tj_all", "target": "TJ", "max effect": 3.0, "activators": ["S1"], "activator Km": [0.3], "activator n": [2.0],
             "activator zone": ["env"]},
            {"name": "tj_na", "target": "TJ", "target ion": "Na", "max effect": 0.5, "inhibitors": ["S3"], "inhibitor Km": [0.2],
             "inhibitor n": [1.0], "inhibitor zone": ["env"]}]
SCENARIOS["mammal_ecm_net_tj"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "internal parameters": {"substances affect Vmem": False},
                    "general network": {"implement network": True, "biomolecules": _ENV_BIO, "reactions": [], "channels": [],
                                        "modulators": _TJ_MODS}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra,
    # the recorder files every change of sim.TJ_modulator between two steps under the scheduled interventions; here the
    # network itself rewrites it inside the loop body — not an input of the step
    drop_sched=("TJ_modulator",))


# a reaction OUTSIDE the cells (write_reactions_env, networks.py:1830-2088; applied at the top of run_loop,
# networks.py:2872-2889): S1 turns into S2 in the bath, inhibited by S3 there
_ENV_RX = [{"name": "S1_to_S2", "reaction zone": "env", "reactants": ["S1"], "reactant multipliers": [1], "Km reactants": [0.2],
            "products": ["S2"], "product multipliers": [1], "Km products": [0.5], "max rate": 5.0e-2, "standard free energy": "None",
            "reaction inhibitors": ["S3"], "inhibitor Km": [0.4], "inhibitor n": [1.0], "inhibitor zone": ["env"]}]
SCENARIOS["mammal_ecm_net_envrx"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "internal parameters": {"substances affect Vmem": False},
                    "general network": {"implement network": True, "biomolecules": _ENV_BIO[:3], "reactions": _ENV_RX, "channels": []}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# ligand-gated channels (Molecule.gating, networks.py:5847-5916): L1 opens a Na/K channel from inside the cell, L2 a Ca
# channel from the extracellular side (it lives in the bath and crosses the membrane slowly)
def _ligand(sub, ions, K, peak, extracell):
    sub["ion channel gating"] = {"channel name": "gate_" + sub["name"], "ion channel target": ions,
                                 "target Hill coefficient": K, "target Hill exponent": 2.0, "peak channel opening": peak,
                                 "acts extracellularly": extracell, "activators": "None", "inhibitors": "None"}
    return sub


_LIG_BIO = [_ligand(_substance("L1", 0.5, Dgj=1e-15, gj_imp=False, cell=0.2, apply_to=["Spot"]), ["Na", "K"], 0.3, 5.0e-17, False),
            _ligand(_env_substance("L2", 1.0e-18, 0, 0.4, 0.0, True, True), ["Ca"], 0.5, 2.0e-17, True)]
SCENARIOS["mammal_ecm_net_lig"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _LIG_BIO, "reactions": [], "channels": []}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# active pumping of substances (Molecule.pump, networks.py:5809-5844): Q1 pumped OUT of the cells with ATP (it also leaks
# back through the membrane), Q2 carried INTO the cells by a facilitated transporter, a cation
def _pumped(sub, into_cell, rate, Km, atp):
    sub["active pumping"] = {"turn on": True, "pump to cell": into_cell, "maximum rate": rate, "pump Km": Km, "uses ATP": atp}
    return sub


_PUMP_BIO = [_pumped(_env_substance("Q1", 1.0e-17, 0, 0.05, 0.5, False, True), False, 2.0e-9, 0.2, True),
             _pumped(_env_substance("Q2", 0.0, 1, 0.4, 0.02, True, False, tj_factor=0.5), True, 1.0e-9, 0.3, False)]
SCENARIOS["mammal_ecm_net_pump"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _PUMP_BIO, "reactions": [], "channels": []}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# transporters (run_loop_transporters, networks.py:2985-3107): a reversible carrier bringing T1 into the cells from the
# bath, and an electrogenic exporter on one tissue profile that moves 3 Na+ out per T2 -> G1 turnover, activated by T1
_TR = [{"name": "carrier", "reaction zone": "cell", "reactants": ["T1"], "reactant multipliers": [1], "Km reactants": [0.5],
        "products": ["T1"], "product multipliers": [1], "Km products": [0.5], "transfered out of cell": [],
        "transfered into cell": ["T1"], "max rate": 5.0e-9, "standard free energy": 0, "apply to": "all", "ignore ECM": True,
        "transporter activators": "None", "transporter inhibitors": "None"},
       {"name": "exporter", "reaction zone": "cell", "reactants": ["Na", "T2"], "reactant multipliers": [3, 1],
        "Km reactants": [10.0, 0.1], "products": ["Na", "G1"], "product multipliers": [3, 1], "Km products": [140.0, 1.0],
        "transfered out of cell": ["Na"], "transfered into cell": [], "max rate": 2.0e-9, "standard free energy": -20e3,
        "apply to": ["Spot"], "ignore ECM": True, "transporter activators": ["T1"], "activator Km": [0.3], "activator n": [1.0],
        "activator zone": ["cell"], "transporter inhibitors": "None"}]
_TR_BIO = [_env_substance("T1", 0.0, 0, 0.5, 0.1, True, True), _substance("T2", 1.0, cell=0.5), _substance("G1", 0.2, cell=0.05)]
SCENARIOS["mammal_ecm_net_trans"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _TR_BIO, "reactions": [], "channels": [],
                                        "transporters": _TR}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# BASELINE configs[3]: the SHIPPED gene regulatory network (extra_configs/grn_basic.yaml: three genes, Hill activation /
# inhibition) run by the second handler (sim.grn.core, sim.py:1305-1319), general network off
SCENARIOS["mammal_ecm_grn"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": False},
                    "gene regulatory network settings": {"gene regulatory network simulated": True}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# 'update intracellular' ON (the default when a config omits the key; the shipped grn_basic.yaml and metabo_basic.yaml do) for
# substances whose membrane value feeds back: Molecule.update_intra relaxes cc_at_mem towards the cell value, with
# electrophoresis in the cell's field for charged ones (networks.py:5714-5806); molecule_mover's membrane and gap-junction
# legs read and move the membrane values (update_at_mems / update_intra branches, sim_toolbox.py:962-1005, 1183-1185).
# (The shipped metabo_basic.yaml itself cannot serve: the reference halts on it in the first step — "ATP in environment
# below zero" — with the default world.)
_INTRA_BIO = [dict(_substance("A", 0.5, Dgj=1e-15, gj_imp=False, cell=0.5, z=-1), **{"Dm": 1.0e-17, "env conc": 0.2,
                                                                                       "update intracellular": True}),
              dict(_substance("B", 0.2, Dgj=1e-14, gj_imp=False, cell=0.05, apply_to=["Spot"]), **{"update intracellular": True}),
              dict(_substance("Cp", 1.0, cell=0.3, z=1), **{"update intracellular": True}),
              _substance("D", 0.3, acts=[("B", 0.1, 1)])]
_INTRA_RX = [{"name": "A_to_B", "reaction zone": "cell", "reactants": ["A"], "reactant multipliers": [1],
              "Km reactants": [0.5], "products": ["B"], "product multipliers": [1], "Km products": [0.1],
              "max rate": 2.0e-2, "standard free energy": "None"}]
SCENARIOS["mammal_ecm_net_intra"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _INTRA_BIO, "reactions": _INTRA_RX,
                                        "channels": []}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# both handlers at once (the reference's own `enable_networks` test scenario, betse_test/_fixture/simconf/simconfwrapper.py:252-259):
# the shipped general network (substance X, Nav1p3 / Kv1p5 / X-inhibited KLeak) AND the shipped gene regulatory network
SCENARIOS["mammal_ecm_net2"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "gene regulatory network settings": {"gene regulatory network simulated": True}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# cell polarizability != 0 TOGETHER with the networks: per-membrane Vmem feeds the channels' gates, the channel
# modulation, the substances' gap-junction and membrane legs and the transporters (vm = sim.vm[m] everywhere)
SCENARIOS["mammal_ecm_polar_net"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "internal parameters": {"cell polarizability": 1.0e-4},
                    "general network": {"implement network": True, "biomolecules": _NET_BIO[3:] + _ENV_BIO[:1] + _TR_BIO[1:],
                                        "reactions": [], "channels": [dict(c) for c in CHANNELS[:2]] + [_NET_CH[2]],
                                        "transporters": [dict(_TR[1], **{"transporter activators": ["S1"]})]}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# regulators OUTSIDE the cells in cell-zone rate laws (get_influencers with zone 'env', networks.py:5242-5265): the gene G1
# is repressed by the bath substance S2 and activated by extracellular K+ at the env square of the cell centre
def _zoned(sub, **zones):
    sub["growth and decay"].update(zones)
    return sub


_ENVZ_BIO = [_env_substance("S2", 0.0, 0, 0.2, 0.0, True, True),
             _zoned(_substance("G1", 2.0, acts=[("K", 5.0, 1)], inh=[("S2", 0.2, 2)]),
                    **{"zone activators": ["env"], "zone inhibitors": ["env"]})]
SCENARIOS["mammal_ecm_net_envzone"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _ENVZ_BIO, "reactions": [], "channels": []}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]}, extra=net_extra)


# the substances' own timed events (networks.py:2929-2933): the bath concentration of B1 is ramped at the global boundary
# (Molecule.update_boundary), C1 is clamped inside the cells for a while (Molecule.cell_clamp_method); both act in the
# first 20 steps of either phase (INIT dt = 1e-2 s, SIM dt = 1e-4 s: two events each)
_EVT_BIO = [dict(_env_substance("B1", 1.0e-17, 0, 0.2, 0.1, True, True),
                 **{"change at bounds": {"event happens": True, "change start": 5.0e-4, "change finish": 0.12,
                                         "change rate": 3.0e-4, "concentration": 1.5}}),
            dict(_substance("C1", 0.5, Dgj=1e-15, gj_imp=False, cell=0.3),
                 **{"clamp cell conc": {"event happens": True, "change start": 3.0e-4, "change finish": 0.05,
                                        "change rate": 2.0e-4, "concentration": 0.9}})]
SCENARIOS["mammal_ecm_net_events"] = dict(
    mods=_m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                    "general network": {"implement network": True, "biomolecules": _EVT_BIO, "reactions": [], "channels": []}}),
    snaps={"init": [1, 2, 5, 13], "sim": [1, 2, 4, 6, 9, 20]}, extra=net_extra)


# dynamic noise (sim.py:1322-1339): a random walk on the protein concentration, one np.random.random(mdl) draw per SIM step
SCENARIOS["mammal_ecm_dynnoise"] = dict(
    mods=_m(NO_NET, SMALL, {"general options": {"ion profile": "mammal"},
                            "variable settings": {"noise": {"dynamic noise": True, "dynamic noise level": 1.0e-6}}}),
    snaps={"init": [1, 2], "sim": [1, 2, 5, 20]})


# The external-voltage event (tissue/event/tisevevolt.py: bound_V ramps, Phi_b = one Dirichlet Poisson solve per step,
# ion_current.py:84-90, subtracted from Vmem in update_V, sim.py:2029) — ramp up, plateau and ramp down inside the
# first 20 SIM steps, left/right electrodes so that it differs from the top/bottom default; ECM and no-ECM
_VOLT = {"apply external voltage": {"event happens": True, "change start": 2.0e-4, "change finish": 1.4e-3,
                                    "change rate": 2.0e-4, "peak voltage": 5.0e-3,
                                    "positive voltage boundary": "left", "negative voltage boundary": "right"}}
SCENARIOS["mammal_ecm_volt"] = dict(mods=_m(NO_NET, SMALL, _VOLT, {"general options": {"ion profile": "mammal"}}),
                                    snaps={"init": [1, 2], "sim": [1, 2, 3, 5, 9, 16, 20]})
SCENARIOS["basic_noecm_volt"] = dict(mods=_m(NO_NET, SMALL, _VOLT, {"apply external voltage": {
    "positive voltage boundary": "top", "negative voltage boundary": "right"},
    "general options": {"simulate extracellular spaces": False}}), snaps={"init": [1, 2], "sim": [1, 3, 9, 20]})


# BASELINE configs[0] as SHIPPED (`betse try`): the unmodified default sim_config.yaml — basic ion profile, ECM on,
# general network on (substance X grown in the 'Spot' profile, Nav1p3 / Kv1p5 / X-inhibited KLeak channels) and the
# cutting event, which removes the 'surgery' wedge at the first SIM step.  Full phases (500 + 350 timesteps) with the
# Vmem / concentration traces at every sampled step: the "final Vmem traces within 1e-6 V" bar of BASELINE.json.
SCENARIOS["default_try"] = dict(mods={}, snaps={"init": [500], "sim": [350]}, extra=net_extra,
                                trace=("vm", "cc_cells"), precut=True, full=True)


# BASELINE configs[0] to the letter: the default sample simulation with the Na/K/Cl/Ca/P(+M) profile and WITHOUT
# extracellular spaces (one well-mixed bath) — network, channels and cutting event as shipped; full phases with traces
SCENARIOS["default_try_noecm"] = dict(mods={"general options": {"ion profile": "mammal", "simulate extracellular spaces": False}},
                                      snaps={"init": [500], "sim": [350]}, extra=net_extra,
                                      trace=("vm", "cc_cells"), precut=True, full=True)


def main(argv):
    import scipy
    names = argv or list(SCENARIOS)
    for name in names:
        sc = SCENARIOS[name]
        cap = refrun.run_reference(sc["mods"], seed=12345, snap_steps=sc["snaps"],
                                   max_steps={"sim": max(max(sc["snaps"]["sim"]), 12)},
                                   extra=sc.get("extra"), tweak_p=sc.get("tweak_p"),
                                   trace=sc.get("trace", ()), precut=sc.get("precut", False),
                                   method=sc.get("method", "_run_sim_core_loop"))
        for f in sc.get("drop_sched", ()):
            for k in [k for k in cap if ".sched." in k and k.endswith("." + f)]:
                cap.pop(k)
        cap["meta.numpy"] = np.array(np.__version__)
        cap["meta.scipy"] = np.array(scipy.__version__)
        cap["meta.seed"] = np.array(12345)
        cap["meta.scenario"] = np.array(name)
        # trim: diagnostics only at the last snapshot of each phase
        for kind in ("init", "sim"):
            ks = sc["snaps"][kind]
            for K in ks[:-1]:
                for f in refrun.DIAG_FIELDS + ["smooth_weight_mem", "smooth_weight_o", "rho_factor",
                                               "Dm_cells", "D_gj", "D_env", "TJ_modulator"]:
                    cap.pop("%s.k%d.%s" % (kind, K, f), None)
        fn = os.path.join(HERE, name + ".npz")
        np.savez_compressed(fn, **cap)
        print("wrote", fn, os.path.getsize(fn) // 1024, "KiB",
              "C=%d M=%d E=%d" % (len(cap["cells.cell_vol"]), len(cap["cells.mem_sa"]),
                                  len(cap["cells.xypts"])))


if __name__ == "__main__":
    main(sys.argv[1:])
