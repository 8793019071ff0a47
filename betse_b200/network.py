"""Host side of the general network / gene regulatory network on the GPU.

``describe_core`` reads a live ``MasterOfNetworks`` (betse/science/chemistry/networks.py) into a
plain *network description*; ``compile_network`` turns a description into the tables
``TissueEngine.set_network`` uploads.  A description is also what the golden fixtures record, so the
GPU tests exercise the same compile path without the reference.

Implemented subset (everything else is refused with the reason, never approximated):

* substances with growth/decay (``write_growth_and_decay``, networks.py:1187-1310) and cell-zone
  reactions (``write_reactions``, 1312-1572), integrated by ``run_loop`` (2805-2914);
* gap-junction transport of substances (``Molecule.transport`` -> ``stb.molecule_mover``,
  sim_toolbox.py:976-1006);
* modulation of voltage-gated channels by substances (``alpha_eval_string``, networks.py:3147-3164).
"""
import numpy as np

from . import ratelaw
from .capi import BetseB200Error


def _is_none(v):
    return v is None or (isinstance(v, str) and v == "None") or (hasattr(v, "__len__") and len(v) == 0)


# run_loop_modulators targets implemented on the device (networks.py:3296-3299); 'TJ' works in the env zone
MOD_TARGETS = {"GJ": 0, "Na/K-ATPase": 1, "TJ": 2}       # TJ: an extracellular-zone rate law rewrites sim.TJ_modulator (networks.py:3301-3317)


def unsupported_reasons(core, p):
    """Why this MasterOfNetworks cannot run on the device (empty list: it can)."""
    bad = []
    if getattr(core, "mit_enabled", False):
        bad.append("mitochondria")
    for name, mod in (getattr(core, "modulators", None) or {}).items():
        if str(mod.target_label) not in MOD_TARGETS:
            bad.append("modulator %r of %s" % (name, mod.target_label))
        if str(mod.target_label) == "TJ" and not bool(getattr(p, "is_ecm", False)):
            bad.append("tight-junction modulator %r without extracellular spaces" % name)
    for name, t in (getattr(core, "transporters", None) or {}).items():
        if str(getattr(t, "reaction_zone", "cell")) != "cell":
            bad.append("transporter %r outside the cell zone" % name)
        if not bool(getattr(p, "is_ecm", False)):
            bad.append("transporter %r without extracellular spaces" % name)
    for attr, what in (("reactions_mit", "mitochondrial reactions"),):
        if len(getattr(core, attr, None) or {}):
            bad.append(what)
    if len(getattr(core, "reactions_env", None) or {}) and not bool(getattr(p, "is_ecm", False)):
        bad.append("extracellular reactions without extracellular spaces")
    for name, m in (getattr(core, "molecules", None) or {}).items():
        why = []
        if bool(getattr(m, "update_intra_conc", False)):
            # Molecule.update_intra with intracellular transport (networks.py:5714-5806): the membrane value becomes state
            # of its own: it relaxes towards the cell value (electrophoresis in the cell's field included), and the
            # membrane and gap-junction legs of molecule_mover read and move IT.  Not implemented where further consumers
            # read it: pumps, ligand gates, transporters (and membrane-zone rate laws, compile_network)
            blockers = []
            if not bool(getattr(p, "is_ecm", False)) and float(getattr(m, "Dm", 0.0) or 0.0) != 0.0:
                blockers.append("membrane-permeable in a tissue without extracellular spaces")
            if bool(getattr(m, "active_pumping", False)) and bool(getattr(m, "use_pumping", False)):
                blockers.append("pumped")
            if bool(getattr(m, "ion_channel_gating", False)) and bool(getattr(m, "use_gating_ligand", False)):
                blockers.append("gating a channel")
            if any(name in list(t.reactants_list) + list(t.products_list) for t in (getattr(core, "transporters", None) or {}).values()):
                blockers.append("moved by a transporter")
            regs = []
            for holder, attrs in (("channels", ("channel_activators_list", "channel_inhibitors_list")),
                                  ("modulators", ("modulator_activators_list", "modulator_inhibitors_list"))):
                for obj in (getattr(core, holder, None) or {}).values():
                    for a in attrs:
                        regs += [str(x) for x in (getattr(obj, a, None) or [])]
            if name in regs:
                blockers.append("regulating a channel or modulator (membrane-zone rate law)")
            if blockers:
                why.append("update intracellular of a substance that is " + ", ".join(blockers))
        if bool(getattr(m, "transmem", False)):
            why.append("transmembrane transport")
        if bool(getattr(m, "change_bounds", False)) and bool(getattr(m, "change_at_bounds", False)) and not bool(getattr(p, "is_ecm", False)):
            why.append("boundary change event without extracellular spaces")
        if bool(getattr(m, "cell_clamp", False)) and bool(getattr(m, "cell_clamp_event", False)) and \
                bool(getattr(m, "active_pumping", False)) and bool(getattr(m, "use_pumping", False)):
            why.append("cell clamp of a pumped substance")
        if _in_env(m) and float(getattr(p, "sharpness", 1.0)) < 1.0:
            why.append("extracellular transport with 'sharpness env' < 1")
        if _in_env(m) and float(getattr(m, "Mu_mem", 0.0) or 0.0) != 0.0:
            why.append("electrophoretic membrane mobility (Mu_mem)")
        if why:
            bad.append("substance %r: %s" % (name, ", ".join(why)))
    for name, r in (getattr(core, "reactions", None) or {}).items():
        if str(getattr(r, "reaction_zone", "cell")) != "cell":
            bad.append("reaction %r outside the cell zone" % name)
    return bad


def _in_env(m):
    """Does the substance cross the membrane or exist outside the cells?  (molecule_mover: the membrane flux needs
    Dm != 0, sim_toolbox.py:943-954; the extracellular transport runs when c_env is not identically 0 or the
    boundary concentration is, sim_toolbox.py:1064-1066)"""
    return (float(getattr(m, "Dm", 0.0) or 0.0) != 0.0 or bool(np.any(np.asarray(m.c_env) != 0.0))
            or float(getattr(m, "c_bound", 0.0) or 0.0) > 1.0e-15)


def pulse(t, t_on, t_off, t_change):
    """tb.pulse (betse/science/math/toolbox.py:353-382): difference of two logistic steps."""
    g = (1 / t_change) * 10
    with np.errstate(over="ignore"):                  # exp -> inf far from the ramps, as in the reference: 1/(1+inf) = 0
        y1 = 1 / (1 + (np.exp(-g * (t - t_on))))
        y2 = 1 / (1 + (np.exp(-g * (t - t_off))))
    return y1 - y2


def event_values(desc, t):
    """Scheduled values of the substances' own events at time ``t`` (host-side scalar logic, like fire_events):
    (c_bound [K], clamp [K] with NaN where no clamp is in force).
    Molecule.update_boundary (networks.py:6043-6066), Molecule.cell_clamp_method (networks.py:6069-6088)."""
    K = len(desc["species"])
    c_bound = np.array(desc.get("c_bound", np.zeros(K)), dtype=float, copy=True)
    clamp = np.full(K, np.nan)
    for ev in desc.get("events", []):
        k = ev["species"]
        if ev.get("bounds") is not None:
            target, start, end, rate, c_envo = ev["bounds"]
            eff = pulse(t, start, end, rate)
            c_bound[k] = target * eff + c_envo * (1 - eff)
        if ev.get("clamp") is not None:
            target, start, end, rate, c_cello = ev["clamp"]
            if start <= t <= end:
                eff = pulse(t, start, end, rate)
                clamp[k] = target * eff + c_cello * (1 - eff)
    return c_bound, clamp


def describe_core(core, sim, p, cells, record_static=True):
    """Live ``MasterOfNetworks`` -> network description (plain dict of strings and arrays)."""
    species = list(core.molecules)
    # short ion names ('Na', 'K', ...: the keys of MasterOfNetworks.cell_concs, networks.py:186-206) in the Simulator's index order
    enabled = [k for k, v in p.ions_dict.items() if v == 1]
    if hasattr(sim, "get_ion"):
        enabled = sorted(enabled, key=lambda k: int(sim.get_ion(k)))
    ions = [str(k) for k in enabled]
    K, C = len(species), len(cells.cell_vol)
    names = list(core.cell_concs.keys())
    rmat = np.asarray(core.reaction_matrix, dtype=float)
    rows = [names.index(s) for s in species]
    other = [i for i in range(len(names)) if i not in rows]
    if rmat.size and np.any(rmat[other] != 0.0):
        raise BetseB200Error("network reactions that produce or consume simulation ions are not implemented")
    desc = {
        "species": species, "ions": ions,
        "c_cells": np.stack([np.asarray(core.molecules[s].c_cells, dtype=float) for s in species]) if K else np.zeros((0, C)),
        "gad_strings": [core.molecules[s].gad_eval_string for s in species],
        "reaction_names": list(core.reactions),
        "reaction_strings": [core.reactions[r].reaction_eval_string for r in core.reactions],
        "stoich": rmat[rows] if rmat.size else np.zeros((K, K)),
        # reactions outside the cells (write_reactions_env, networks.py:1830-2088; applied at the top of run_loop,
        # networks.py:2872-2889): extracellular-zone rate laws, env_concs += reaction_matrix_env . rates * dt
        "reaction_env_names": list(getattr(core, "reactions_env", None) or {}),
        "reaction_env_strings": [r.reaction_eval_string for r in (getattr(core, "reactions_env", None) or {}).values()],
        "growth_targets": [np.asarray(core.molecules[s].growth_targets_cell, dtype=np.int64) for s in species],
        "Dgj": np.array([-1.0 if core.molecules[s].ignoreGJ else float(core.molecules[s].Dgj) for s in species]),
        "z": np.array([float(core.molecules[s].z) for s in species]),
        "time_factor": np.array([float(core.molecules[s].modify_time_factor) for s in species]),
        "scale_factor": np.array([float(getattr(core.molecules[s], "scale_factor", 1.0)) for s in species]),
        "chan_names": list(core.channels),
        "chan_mod_strings": [core.channels[c].alpha_eval_string for c in core.channels],
        # run_loop_modulators (networks.py:3282-3325): modulator = max_val * eval(alpha_eval_string) -> sim.gj_block / NaKATP_block
        "modulator_names": list(getattr(core, "modulators", None) or {}),
        "modulator_strings": [m.alpha_eval_string for m in (getattr(core, "modulators", None) or {}).values()],
        "modulator_targets": [str(m.target_label) for m in (getattr(core, "modulators", None) or {}).values()],
        "modulator_max": np.array([float(m.max_val) for m in (getattr(core, "modulators", None) or {}).values()]),
        # tight-junction modulators: the ion they act on (Modulator.init_modulator, networks.py:6684-6692; -1 = all ions) and
        # the env squares of the barrier (sim.TJ_targets, sim.py:2394-2395)
        "modulator_ions": np.array([-1 if getattr(m, "ion_i", None) is None else int(m.ion_i)
                                    for m in (getattr(core, "modulators", None) or {}).values()], dtype=np.int64),
        "tj_targets": np.asarray(getattr(sim, "TJ_targets", np.zeros(0)), dtype=np.int64)
        if any(str(m.target_label) == "TJ" for m in (getattr(core, "modulators", None) or {}).values()) else np.zeros(0, dtype=np.int64),
        "n_env": int(len(cells.xypts)) if getattr(cells, "xypts", None) is not None else 0,
        "static": {},
    }
    if desc["reaction_env_names"]:
        names_e = list(core.env_concs.keys())                     # rows of reaction_matrix_env (create_reaction_matrix_env, networks.py:2758-2790)
        rme = np.asarray(core.reaction_matrix_env, dtype=float)
        rows_e = [names_e.index(s) for s in species]
        other_e = [i for i in range(len(names_e)) if i not in rows_e]
        if np.any(rme[other_e] != 0.0):
            raise BetseB200Error("extracellular reactions that produce or consume simulation ions are not implemented")
        desc["stoich_env"] = rme[rows_e]
    # Molecule.transport -> stb.molecule_mover (networks.py:5670-5700, sim_toolbox.py:909-1153): membrane and
    # extracellular legs of the substances that have them
    events = []
    for k, s in enumerate(species):
        m = core.molecules[s]
        ev = {"species": k, "bounds": None, "clamp": None}
        if bool(getattr(m, "change_bounds", False)) and bool(getattr(m, "change_at_bounds", False)):
            ev["bounds"] = (float(m.change_bounds_target), float(m.change_bounds_start), float(m.change_bounds_end),
                            float(m.change_bounds_rate), float(m.c_envo))
        if bool(getattr(m, "cell_clamp", False)) and bool(getattr(m, "cell_clamp_event", False)):
            ev["clamp"] = (float(m.cell_clamp_target), float(m.cell_clamp_start), float(m.cell_clamp_end),
                           float(m.cell_clamp_rate), float(m.c_cello))
        if ev["bounds"] is not None or ev["clamp"] is not None:
            events.append(ev)
    if events:
        desc["events"] = events
    intra = np.array([bool(getattr(core.molecules[s], "update_intra_conc", False)) for s in species], dtype=np.uint8)
    if intra.any():
        desc.update({"intra_on": intra, "Do": np.array([float(core.molecules[s].Do or 0.0) for s in species]),
                     "mu_mem": np.array([float(getattr(core.molecules[s], "Mu_mem", 0.0) or 0.0) for s in species]),
                     "c_mems": np.stack([np.asarray(core.molecules[s].cc_at_mem, dtype=float) * np.ones(len(cells.mem_sa)) for s in species])})
    # run_loop_transporters (networks.py:2985-3107): flux = rho_pump * eval(transporter_eval_string) on every membrane;
    # each reactant / product moves by coeff * (-/+) sum_mems(flux*mem_sa)/cell_vol in the cells ('mem_concs' tag) or
    # coeff * div_env(-/+flux) outside ('env_concs' tag), on the transporter's target cells / env squares
    trans = []
    for name, t in (getattr(core, "transporters", None) or {}).items():
        terms = [(str(x), float(c), str(tag), -1) for x, c, tag in zip(t.reactants_list, t.reactants_coeff, t.react_transport_tag)]
        terms += [(str(x), float(c), str(tag), +1) for x, c, tag in zip(t.products_list, t.products_coeff, t.prod_transport_tag)]
        for x, c, tag, sg in terms:
            if tag not in ("mem_concs", "env_concs"):
                raise BetseB200Error("transporter %r: zone %r is not implemented" % (name, tag))
        trans.append({"name": str(name), "eval_string": str(t.transporter_eval_string), "net_z": float(t.net_z), "terms": terms,
                      "targets_cell": np.asarray(t.transporter_targets_cell, dtype=np.int64),
                      "targets_mem": np.asarray(t.transporter_targets_mem, dtype=np.int64),
                      "targets_env": np.asarray(t.transporter_targets_env, dtype=np.int64)})
    if trans:
        desc["transporters"] = trans
    # Molecule.gating (networks.py:5847-5916): ligand-gated channels, one entry per (substance, conducted ion)
    lig = []
    for k, s in enumerate(species):
        m = core.molecules[s]
        if bool(getattr(m, "ion_channel_gating", False)) and bool(getattr(m, "use_gating_ligand", False)):
            for ion_tag in m.gating_ion:
                lig.append({"species": k, "ion": int(ion_tag), "K": float(m.gating_Hill_K), "n": float(m.gating_Hill_n),
                            "max": float(m.gating_max_val), "extracell": bool(m.gating_extracell),
                            "mod_string": str(m.gating_mod_eval_string)})
    if lig:
        desc["ligand_gates"] = lig
    # Molecule.pump (networks.py:5809-5844): active pump (stb.molecule_pump) or facilitated transporter
    # (stb.molecule_transporter) of the substance itself, applied right after its growth/decay
    pumps = [{"species": k, "into_cell": bool(core.molecules[s].pump_to_cell), "max": float(core.molecules[s].pump_max_val),
              "Km": float(core.molecules[s].pump_Km), "uses_ATP": bool(core.molecules[s].pumps_use_ATP)}
             for k, s in enumerate(species)
             if bool(getattr(core.molecules[s], "active_pumping", False)) and bool(getattr(core.molecules[s], "use_pumping", False))]
    if pumps:
        desc["pumps"] = pumps
        if any(g["extracell"] and any(q["species"] == g["species"] for q in pumps) for g in lig):
            raise BetseB200Error("a pumped substance that also gates a channel from outside the cell is not implemented")
    env_on = np.array([_in_env(core.molecules[s]) or any(g["species"] == k and g["extracell"] for g in lig)
                       or any(ev["species"] == k and ev["bounds"] is not None for ev in events)
                       or any(q["species"] == k for q in pumps)
                       or any(x == s and tag == "env_concs" for t in trans for x, _, tag, _ in t["terms"])   # moved outside by a transporter
                       for k, s in enumerate(species)], dtype=np.uint8)
    if env_on.any() and bool(getattr(p, "is_ecm", False)):
        E = int(np.asarray(sim.D_env_weight).size)
        D_env = np.zeros((K, E))
        c_env = np.zeros((K, E))
        for k, s in enumerate(species):
            m = core.molecules[s]
            if not env_on[k]:
                continue
            mult = np.ones(E)
            if not bool(m.ignoreTJ):                    # sim_toolbox.py:1079-1085
                mult = np.array(sim.D_env_weight, dtype=float).ravel().copy()
                tj = np.asarray(sim.TJ_targets, dtype=np.int64)
                mult[tj] = mult[tj] * float(m.TJ_factor)
            D_env[k] = mult * float(m.Do)
            c_env[k] = np.asarray(m.c_env, dtype=float)
        desc.update({"env_on": env_on, "Dm": np.array([float(core.molecules[s].Dm or 0.0) for s in species]),
                     "c_bound": np.array([float(core.molecules[s].c_bound or 0.0) for s in species]),
                     "c_env": c_env, "D_env": D_env})
    if record_static:
        # resolve every static leaf once so that the description is self-contained
        rec = {}
        compile_network(desc, C, len(cells.mem_sa), ratelaw.live_resolver(core, sim, p, cells, record=rec))
        desc["static"] = rec
    return desc


def compile_network(desc, n_cells, n_mems, resolver=None):
    """Description -> {'species','tables','rate_programs','mod_programs','mod_index', ...} for
    ``TissueEngine.set_network``.  ``resolver`` defaults to the description's recorded static values."""
    if resolver is None:
        resolver = ratelaw.table_resolver(desc["static"])
    species = list(desc["species"])
    K = len(species)
    tabs = ratelaw.Tables(species, list(desc["ions"]), n_cells, n_mems, int(desc.get("n_env", 0)))
    mtargets = list(desc.get("modulator_targets", []))
    try:
        rates = [ratelaw.compile_expr(s, tabs, resolver, "cell") for s in desc["gad_strings"]]
        rates += [ratelaw.compile_expr(s, tabs, resolver, "cell") for s in desc["reaction_strings"]]
        mods = [ratelaw.compile_expr(s, tabs, resolver, "mem") for s in desc["chan_mod_strings"]]
        modulators = [ratelaw.compile_expr(s, tabs, resolver, "env" if t == "TJ" else "mem")
                      for s, t in zip(desc.get("modulator_strings", []), mtargets)]
        env_rx = [ratelaw.compile_expr(s, tabs, resolver, "env") for s in desc.get("reaction_env_strings", [])]
    except ratelaw.RateLawError as e:
        raise BetseB200Error("network rate law not supported on the device: %s" % e)
    stoich = np.asarray(desc["stoich"], dtype=float).reshape(K, -1)
    if stoich.shape[1] != len(rates):
        raise BetseB200Error("reaction_matrix has %d columns for %d rate laws" % (stoich.shape[1], len(rates)))
    mask = np.zeros((K, n_cells), dtype=np.uint8)
    for k, tg in enumerate(desc["growth_targets"]):
        mask[k, np.asarray(tg, dtype=np.int64)] = 1
    # modulators that fold to the constant 1 need no program (moddy == 1)
    mod_index, mod_programs = [], []
    for pr in mods:
        if len(pr.code) == 1 and pr.code[0][0] == ratelaw.PUSHC and tabs.consts[pr.code[0][1]] == 1.0:
            mod_index.append(-1)
        else:
            mod_index.append(len(rates) + len(mod_programs))
            mod_programs.append(pr)
    # sim modulators: always a program (a constant one when nothing regulates them), after the channel programs
    modulator_index = []
    for pr in modulators:
        modulator_index.append(len(rates) + len(mod_programs))
        mod_programs.append(pr)
    for t in desc.get("modulator_targets", []):
        if t not in MOD_TARGETS:
            raise BetseB200Error("modulator target %r is not implemented" % t)
    # extracellular reactions: their programs follow the modulators'
    env_rx_index = []
    for pr in env_rx:
        env_rx_index.append(len(rates) + len(mod_programs))
        mod_programs.append(pr)
    stoich_env = np.asarray(desc.get("stoich_env", np.zeros((K, 0))), dtype=float).reshape(K, -1)
    if stoich_env.shape[1] != len(env_rx):
        raise BetseB200Error("reaction_matrix_env has %d columns for %d extracellular reactions" % (stoich_env.shape[1], len(env_rx)))
    if env_rx:
        # run_loop applies them before anything else of the step reads the env concentrations (networks.py:2872-2889);
        # the device evaluates the other zones' rate laws later in the step: a substance they move must not be read there
        moved = {k for k in range(K) if np.any(stoich_env[k] != 0.0)}
        eo_ = np.asarray(desc.get("env_on", np.zeros(K)), dtype=bool)
        for k in moved:
            if not eo_[k]:
                raise BetseB200Error("an extracellular reaction moves %r, which does not exist outside the cells" % species[k])
        for pr in rates:            # (channel and modulator programs run before run_loop, like the reference's)
            for op, arg in pr.code:
                if op == ratelaw.PUSHE and arg in moved:
                    raise BetseB200Error("a cell-zone rate law reads %r outside the cells while an extracellular reaction "
                                         "moves it: not implemented" % species[arg])
    # transporters: one membrane-zone program each, after the modulators' programs
    transporters = []
    ions = list(desc["ions"])
    for t in desc.get("transporters", []):
        try:
            pr = ratelaw.compile_expr(t["eval_string"], tabs, resolver, "mem")
        except ratelaw.RateLawError as e:
            raise BetseB200Error("transporter %r not supported on the device: %s" % (t["name"], e))
        terms = []
        for x, coeff, tag, sign in t["terms"]:
            env = tag == "env_concs"
            if x in species:
                terms.append((3 if env else 2, species.index(x), int(sign), float(coeff)))
            elif x in ions:
                terms.append((1 if env else 0, ions.index(x), int(sign), float(coeff)))
            else:
                raise BetseB200Error("transporter %r moves %r, which is neither a substance nor a simulated ion" % (t["name"], x))
        cm = np.zeros(n_cells, dtype=np.uint8)
        cm[np.asarray(t["targets_cell"], dtype=np.int64)] = 1
        em = None
        if len(t["targets_env"]):
            n_env = int(np.asarray(desc["c_env"]).shape[1]) if "c_env" in desc else int(np.max(t["targets_env"])) + 1
            em = np.zeros(n_env, dtype=np.uint8)
            em[np.asarray(t["targets_env"], dtype=np.int64)] = 1
        mm = np.zeros(n_mems, dtype=np.uint8)
        mm[np.asarray(t["targets_mem"], dtype=np.int64)] = 1
        transporters.append({"prog": len(rates) + len(mod_programs), "net_z": float(t["net_z"]), "terms": terms,
                             "cell_mask": None if cm.all() else cm, "env_mask": em, "mem_mask": None if mm.all() else mm})
        mod_programs.append(pr)
    io = np.asarray(desc.get("intra_on", np.zeros(K)), dtype=bool)
    for pr in mods + modulators:
        for op, arg in pr.code:
            if op == ratelaw.PUSHS and io[arg]:
                raise BetseB200Error("a membrane-zone rate law reads %r, whose membrane value is transported inside the cell "
                                     "('update intracellular'): not implemented" % species[arg])
    eo = np.asarray(desc.get("env_on", np.zeros(K)), dtype=bool)
    for k in sorted(tabs.env_species):
        if not eo[k]:
            raise BetseB200Error("a rate law reads %r outside the cells, where it does not exist" % species[k])
    # ligand-gated channels: the regulation of the gate itself (gating_mod_eval_string) must fold to a constant — it
    # is evaluated in the middle of run_loop's per-substance sequence (networks.py:2922-2925), which a device kernel
    # working from the step's starting concentrations cannot mirror for other substances
    gates = []
    for g in desc.get("ligand_gates", []):
        try:
            pr = ratelaw.compile_expr(g["mod_string"], tabs, resolver, "mem")
        except ratelaw.RateLawError as e:
            raise BetseB200Error("ligand-gated channel: %s" % e)
        if len(pr.code) != 1 or pr.code[0][0] != ratelaw.PUSHC:
            raise BetseB200Error("ligand-gated channels regulated by further substances are not implemented")
        gates.append(dict(g, mod=float(tabs.consts[pr.code[0][1]])))
    return {"species": species, "tables": tabs, "rate_programs": rates, "mod_programs": mod_programs,
            "mod_index": mod_index, "ligand_gates": gates, "pumps": list(desc.get("pumps", [])), "transporters": transporters,
            "events": list(desc.get("events", [])),
            "modulators": [(MOD_TARGETS[t], i, float(mx), int(ion)) for t, i, mx, ion in
                           zip(mtargets, modulator_index, desc.get("modulator_max", []),
                               desc.get("modulator_ions", [-1] * len(mtargets)))],
            "tj_targets": np.asarray(desc.get("tj_targets", np.zeros(0)), dtype=np.int64),
            "env_rx_index": env_rx_index, "stoich_env": stoich_env, "c_cells": np.asarray(desc["c_cells"], dtype=float), "stoich": stoich,
            "growth_mask": None if mask.all() else mask, "Dgj": np.asarray(desc["Dgj"], dtype=float),
            "z": np.asarray(desc["z"], dtype=float), "time_factor": np.asarray(desc["time_factor"], dtype=float),
            "chan_names": list(desc["chan_names"]),
            **{k: np.asarray(desc[k]) for k in ("env_on", "Dm", "c_bound", "c_env", "D_env", "scale_factor", "intra_on", "Do", "mu_mem", "c_mems") if k in desc}}


# ---- flat (npz-friendly) form of a description, used by the golden fixtures
def flatten(desc, prefix):
    out = {prefix + "species": np.array(desc["species"]), prefix + "ions": np.array(desc["ions"]),
           prefix + "c_cells": np.asarray(desc["c_cells"]), prefix + "gad_strings": np.array(desc["gad_strings"]),
           prefix + "reaction_names": np.array(desc["reaction_names"], dtype=str),
           prefix + "reaction_strings": np.array(desc["reaction_strings"], dtype=str),
           prefix + "stoich": np.asarray(desc["stoich"]), prefix + "Dgj": desc["Dgj"], prefix + "z": desc["z"],
           prefix + "time_factor": desc["time_factor"], prefix + "chan_names": np.array(desc["chan_names"], dtype=str),
           prefix + "chan_mod_strings": np.array(desc["chan_mod_strings"], dtype=str),
           prefix + "static_keys": np.array(list(desc["static"].keys()), dtype=str)}
    if desc.get("reaction_env_names"):
        out.update({prefix + "reaction_env_names": np.array(desc["reaction_env_names"], dtype=str),
                    prefix + "reaction_env_strings": np.array(desc["reaction_env_strings"], dtype=str),
                    prefix + "stoich_env": np.asarray(desc["stoich_env"], dtype=float),
                    prefix + "n_env": np.asarray(int(desc.get("n_env", 0)))})
    for j, g in enumerate(desc.get("ligand_gates", [])):
        out["%slig%d" % (prefix, j)] = np.array([g["species"], g["ion"], g["K"], g["n"], g["max"], float(g["extracell"])])
        out["%slig%d.mod_string" % (prefix, j)] = np.array(g["mod_string"])
    for j, t in enumerate(desc.get("transporters", [])):
        pre = "%strans%d." % (prefix, j)
        out.update({pre + "name": np.array(t["name"]), pre + "eval_string": np.array(t["eval_string"]), pre + "net_z": np.asarray(t["net_z"]),
                    pre + "term_names": np.array([x for x, _, _, _ in t["terms"]], dtype=str),
                    pre + "term_tags": np.array([tag for _, _, tag, _ in t["terms"]], dtype=str),
                    pre + "term_coeff": np.array([c for _, c, _, _ in t["terms"]], dtype=float),
                    pre + "term_sign": np.array([sg for _, _, _, sg in t["terms"]], dtype=np.int64),
                    pre + "targets_cell": np.asarray(t["targets_cell"], dtype=np.int64),
                    pre + "targets_mem": np.asarray(t["targets_mem"], dtype=np.int64),
                    pre + "targets_env": np.asarray(t["targets_env"], dtype=np.int64)})
    for j, ev in enumerate(desc.get("events", [])):
        nan5 = (np.nan,) * 5
        out["%sevent%d" % (prefix, j)] = np.array((ev["species"],) + tuple(ev["bounds"] or nan5) + tuple(ev["clamp"] or nan5), dtype=float)
    for j, q in enumerate(desc.get("pumps", [])):
        out["%spump%d" % (prefix, j)] = np.array([q["species"], float(q["into_cell"]), q["max"], q["Km"], float(q["uses_ATP"])])
    if desc.get("modulator_names"):
        out.update({prefix + "modulator_names": np.array(desc["modulator_names"], dtype=str),
                    prefix + "modulator_strings": np.array(desc["modulator_strings"], dtype=str),
                    prefix + "modulator_targets": np.array(desc["modulator_targets"], dtype=str),
                    prefix + "modulator_max": np.asarray(desc["modulator_max"], dtype=float),
                    prefix + "modulator_ions": np.asarray(desc.get("modulator_ions", -np.ones(len(desc["modulator_names"]))), dtype=np.int64),
                    prefix + "tj_targets": np.asarray(desc.get("tj_targets", np.zeros(0)), dtype=np.int64),
                    prefix + "n_env": np.asarray(int(desc.get("n_env", 0)))})
    for k in ("env_on", "Dm", "c_bound", "c_env", "D_env", "scale_factor", "intra_on", "Do", "mu_mem", "c_mems"):
        if k in desc:
            out[prefix + k] = np.asarray(desc[k])
    for k, tg in enumerate(desc["growth_targets"]):
        out["%sgrowth_targets%d" % (prefix, k)] = np.asarray(tg, dtype=np.int64)
    for j, v in enumerate(desc["static"].values()):
        out["%sstatic%d" % (prefix, j)] = np.asarray(v, dtype=np.float64)
    return out


def unflatten(cap, prefix):
    g = lambda k: cap[prefix + k]
    species = [str(x) for x in g("species")]
    keys = [str(x) for x in g("static_keys")]
    mods = {}
    j = 0
    while "%slig%d" % (prefix, j) in cap:
        v = np.asarray(cap["%slig%d" % (prefix, j)], dtype=float)
        mods.setdefault("ligand_gates", []).append({"species": int(v[0]), "ion": int(v[1]), "K": float(v[2]), "n": float(v[3]),
                                                    "max": float(v[4]), "extracell": bool(v[5]),
                                                    "mod_string": str(cap["%slig%d.mod_string" % (prefix, j)])})
        j += 1
    j = 0
    while "%strans%d.name" % (prefix, j) in cap:
        pre = "%strans%d." % (prefix, j)
        mods.setdefault("transporters", []).append({
            "name": str(cap[pre + "name"]), "eval_string": str(cap[pre + "eval_string"]), "net_z": float(cap[pre + "net_z"]),
            "terms": [(str(x), float(c), str(tag), int(sg)) for x, c, tag, sg in
                      zip(cap[pre + "term_names"], cap[pre + "term_coeff"], cap[pre + "term_tags"], cap[pre + "term_sign"])],
            "targets_cell": np.asarray(cap[pre + "targets_cell"]), "targets_mem": np.asarray(cap[pre + "targets_mem"]),
            "targets_env": np.asarray(cap[pre + "targets_env"])})
        j += 1
    j = 0
    while "%sevent%d" % (prefix, j) in cap:
        v = np.asarray(cap["%sevent%d" % (prefix, j)], dtype=float)
        mods.setdefault("events", []).append({"species": int(v[0]), "bounds": None if np.isnan(v[1]) else tuple(float(x) for x in v[1:6]),
                                              "clamp": None if np.isnan(v[6]) else tuple(float(x) for x in v[6:11])})
        j += 1
    j = 0
    while "%spump%d" % (prefix, j) in cap:
        v = np.asarray(cap["%spump%d" % (prefix, j)], dtype=float)
        mods.setdefault("pumps", []).append({"species": int(v[0]), "into_cell": bool(v[1]), "max": float(v[2]), "Km": float(v[3]),
                                             "uses_ATP": bool(v[4])})
        j += 1
    if prefix + "reaction_env_names" in cap:
        mods = {**mods, **{"reaction_env_names": [str(x) for x in g("reaction_env_names")],
                           "reaction_env_strings": [str(x) for x in g("reaction_env_strings")],
                           "stoich_env": np.asarray(g("stoich_env")), "n_env": int(g("n_env"))}}
    if prefix + "modulator_names" in cap:
        mods = {**mods, **{"modulator_names": [str(x) for x in g("modulator_names")], "modulator_strings": [str(x) for x in g("modulator_strings")],
                "modulator_targets": [str(x) for x in g("modulator_targets")], "modulator_max": np.asarray(g("modulator_max"))}}
        if prefix + "modulator_ions" in cap:
            mods.update({"modulator_ions": np.asarray(g("modulator_ions")), "tj_targets": np.asarray(g("tj_targets")),
                         "n_env": int(g("n_env"))})
    return {**mods, **{k: np.asarray(cap[prefix + k]) for k in ("env_on", "Dm", "c_bound", "c_env", "D_env", "scale_factor", "intra_on", "Do", "mu_mem", "c_mems") if prefix + k in cap},
            "species": species, "ions": [str(x) for x in g("ions")], "c_cells": np.asarray(g("c_cells")),
            "gad_strings": [str(x) for x in g("gad_strings")],
            "reaction_names": [str(x) for x in g("reaction_names")],
            "reaction_strings": [str(x) for x in g("reaction_strings")], "stoich": np.asarray(g("stoich")),
            "growth_targets": [np.asarray(cap["%sgrowth_targets%d" % (prefix, k)]) for k in range(len(species))],
            "Dgj": np.asarray(g("Dgj")), "z": np.asarray(g("z")), "time_factor": np.asarray(g("time_factor")),
            "chan_names": [str(x) for x in g("chan_names")],
            "chan_mod_strings": [str(x) for x in g("chan_mod_strings")],
            "static": {k: (float(cap["%sstatic%d" % (prefix, j)]) if np.ndim(cap["%sstatic%d" % (prefix, j)]) == 0
                           else np.asarray(cap["%sstatic%d" % (prefix, j)])) for j, k in enumerate(keys)}}
