"""General network on the GPU (BASELINE configs[3]) through the C ABI: growth/decay with
Hill regulation, a cell-zone reaction, gap-junction transport of a substance and substance-modulated K channels,
against the REAL reference's recorded run (tests/golden/mammal_ecm_net.npz) and against the oracle on a
synthetic 50 k-cell tissue."""
import numpy as np
import pytest

from betse_b200.network import event_values as netlib_events
from tests import util

pytestmark = pytest.mark.gpu


def _attach(eng, desc, specs, phase_init):
    from betse_b200 import network as netlib
    comp = netlib.compile_network(desc, eng.Co, eng.M)
    eng.set_network(comp, handler=0)
    for c in specs:
        c["handler"] = 0
        c["mod_prog"] = comp["mod_index"][comp["chan_names"].index(c["name"])]
    eng.set_channels(specs, phase_init=phase_init)
    return comp


@pytest.mark.parametrize("kind", ["init", "sim"])
# _env: membrane + extracellular transport of substances; _envq: charged substances with 'substances affect Vmem' on
# _mod: sim modulators (gap junctions, Na/K-ATPase); _lig: ligand-gated channels (intracellular and extracellular ligand)
@pytest.mark.parametrize("fixture", ["mammal_ecm_net", "mammal_ecm_net_env", "mammal_ecm_net_envq", "mammal_ecm_net_mod",
                                     "mammal_ecm_net_lig", "mammal_ecm_net_pump",    # _pump: Molecule.pump (ATP pump out, carrier in)
                                     "mammal_ecm_net_trans",                         # _trans: transporters (carrier; electrogenic exporter)
                                     "mammal_ecm_polar_net",                         # per-membrane Vmem (polarizability) under all of it
                                     "mammal_ecm_net_envzone",                       # cell-zone rate laws regulated from outside the cells
                                     "mammal_ecm_net_events",                        # boundary ramp and cell clamp of substances
                                     "mammal_ecm_net_intra",                         # 'update intracellular': transported membrane values
                                     "mammal_ecm_net_tj",                            # tight-junction modulators (extracellular-zone rate laws)
                                     "mammal_ecm_net_envrx"])                        # a reaction outside the cells (write_reactions_env)
def test_network_matches_reference(fixture, kind):
    from betse_b200.engine import TissueEngine
    cap = util.load_golden(fixture)
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
    specs = util.channels_of(cap, kind)
    desc = util.networks_of(cap, kind)[0]
    _attach(eng, desc, specs, kind == "init")
    active = [c for c in specs if not (kind == "init" and not c["init_active"])]
    n = 0
    snaps = util.snap_steps(cap, kind)
    for K in snaps:
        last = K == snaps[-1]
        while n < K:
            util.apply_schedule(eng, cap, kind, n + 1)
            if desc.get("events"):          # the substances' own timed events: the host evaluates the schedule of the step's time
                eng.set_network_events(0, *netlib_events(desc, float(cap[kind + ".time_steps"][n])))
            st = eng.step(1, diag=(last and n + 1 == K))
            assert not (st & (3 | 16)), st
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        extra = ["Jmem", "Jgj", "Jn", "I_mem", "Jtx", "Jty"] if last else []      # carry extra_J_mem / extra_Jenv (ion_current.py:27, 53-54)
        got = eng.download([f for f in list(util.STATE) + util.ENV_STATE + extra if f in ref])
        tols = util.gpu_tolerances(cap, kind, ref)
        for f, a in got.items():
            err = float(np.max(np.abs(np.asarray(a).reshape(np.shape(ref[f])) - ref[f])))
            assert err <= tols[f], (kind, K, f, err, tols[f])
        c, rates = eng.network_state(0, rates=True)
        want = ref["net0.c_cells"]
        for k, nme in enumerate(desc["species"]):
            assert util.rel_err(c[k], want[k]) <= 1e-10, (kind, K, nme, util.rel_err(c[k], want[k]))
        if "net0.c_env" in ref:
            ce = eng.network_env_state(0)
            for k, nme in enumerate(desc["species"]):
                if desc["env_on"][k]:
                    assert util.rel_err(ce[k], ref["net0.c_env"][k]) <= 1e-10, (kind, K, nme, "env", util.rel_err(ce[k], ref["net0.c_env"][k]))
        if "net0.reaction_rates" in ref:
            rr = ref["net0.reaction_rates"]
            assert util.rel_err(rates[-rr.shape[0]:], rr) <= 1e-10
        if "net0.c_mems" in ref and np.any(desc.get("intra_on", 0)):
            cm = eng.network_mem_state(0)
            for k, nme in enumerate(desc["species"]):
                assert util.rel_err(cm[k], ref["net0.c_mems"][k]) <= 1e-10, (kind, K, nme, "mems", util.rel_err(cm[k], ref["net0.c_mems"][k]))
        if fixture == "mammal_ecm_net_tj" and "TJ_modulator" in ref:
            tjm = eng.tj_modulator()
            assert util.rel_err(tjm, ref["TJ_modulator"].reshape(tjm.shape)) <= 1e-10, (kind, K, "TJ_modulator")
        for k, ch in enumerate(active):
            j = [s["name"] for s in specs].index(ch["name"])
            stt = eng.channel_state(k)
            r = ref["chan%d.DChan" % j]
            assert np.max(np.abs(stt["DChan"] - r)) <= 1e-10 * max(np.max(np.abs(r)), 1e-300), (kind, K, ch["name"])
    eng.close()


def test_network_vs_oracle_synthetic_50k():
    """BASELINE configs[3]: 50 k-cell tissue with a per-cell regulatory network coupled to Vmem (the strings and
    tables of the recorded network, re-targeted to the synthetic tissue), 12 steps."""
    from betse_b200 import channels as chlib
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    cap = util.load_golden("mammal_ecm_net")
    desc = util.networks_of(cap, "sim")[0]
    mesh, p, st = synth.make_tissue(50_000)
    C, M = len(mesh["cell_vol"]), len(mesh["mem_sa"])
    rng = np.random.default_rng(3)
    desc = dict(desc)
    desc["c_cells"] = rng.uniform(0.02, 1.5, (len(desc["species"]), C))
    desc["growth_targets"] = [np.arange(C), np.arange(C), np.arange(C), np.arange(0, C, 7)]
    desc["static"] = {k: (v if np.ndim(v) == 0 else np.ones(M if "mdl" in k else C)) for k, v in desc["static"].items()}
    desc["static"]["self.molecules['G1'].growth_mod_function_cells"] = rng.uniform(0.5, 1.5, C)   # a spatial modulator function
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    ora0 = OracleSim(mesh, p, st)
    ora0.diagnostics = False
    ora0.update_V()
    specs = []
    for name, model, dm in (("Nav", "Nav1p3", 2.0e-14), ("Kv", "Kv1p5", 1.0e-15), ("K_Leak", "KLeak", 0.6e-17),
                            ("Cav", "Cav1p2", 1.0e-15)):
        m0, h0 = chlib.initial_state(model, ora0.vm)
        specs.append(chlib.make_channel(name, model, dm, m=m0, h=h0))
    ora = OracleSim(mesh, p, st, channels=[dict(s) for s in specs], networks=[desc])
    ora.diagnostics = False
    ora.update_V()
    _attach(eng, desc, specs, False)
    for n in range(12):
        s = eng.step(1)
        ora.step()
        assert not (s & (3 | 16))
    got = eng.download(["cc_cells", "cc_env", "vm", "gjopen"])
    for f, a in got.items():
        r = np.asarray(getattr(ora, f))
        scale = max(float(np.max(np.abs(r))), 1e-300)
        tol = 1e-10 * scale if f != "vm" else max(1e-10 * scale, 4e-12)
        assert float(np.max(np.abs(a.reshape(r.shape) - r))) <= tol, f
    c = eng.network_state(0)
    for k, nme in enumerate(desc["species"]):
        assert util.rel_err(c[k], ora.networks[0].c[nme]) <= 1e-10, nme
    eng.close()


@pytest.mark.parametrize("kind", ["init", "sim"])
def test_gene_network_handler_matches_reference(kind):
    """BASELINE configs[3]: the SHIPPED gene regulatory network (extra_configs/grn_basic.yaml) run by the second handler
    (sim.grn.core, sim.py:1305-1319) — betse_set_network(handler = 1) — against the real reference's recorded run."""
    from betse_b200 import network as netlib
    from betse_b200.engine import TissueEngine
    cap = util.load_golden("mammal_ecm_grn")
    assert util.network_handlers(cap, kind) == [1]
    desc = util.networks_of(cap, kind)[0]
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
    eng.set_network(netlib.compile_network(desc, eng.Co, eng.M), handler=1)
    eng.set_channels([], phase_init=(kind == "init"))
    n = 0
    for K in util.snap_steps(cap, kind):
        while n < K:
            util.apply_schedule(eng, cap, kind, n + 1)
            assert not (eng.step(1) & (3 | 16))
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        got = eng.download([f for f in list(util.STATE) + util.ENV_STATE if f in ref])
        tols = util.gpu_tolerances(cap, kind, ref)
        for f, a in got.items():
            err = float(np.max(np.abs(np.asarray(a).reshape(np.shape(ref[f])) - ref[f])))
            assert err <= tols[f], (kind, K, f, err, tols[f])
        c = eng.network_state(1)
        cm = eng.network_mem_state(1)             # grn_basic.yaml leaves 'update intracellular' at its default (on)
        for k, nme in enumerate(desc["species"]):
            assert util.rel_err(c[k], ref["net1.c_cells"][k]) <= 1e-10, (kind, K, nme)
            assert util.rel_err(cm[k], ref["net1.c_mems"][k]) <= 1e-10, (kind, K, nme, "mems")
    eng.close()


@pytest.mark.parametrize("kind", ["init", "sim"])
def test_both_handlers_match_reference(kind):
    """The reference's `enable_networks` scenario: the shipped general network (handler 0: substance X, three channels, one
    of them inhibited by X) and the shipped gene regulatory network (handler 1) in the same step (sim.py:1290-1319)."""
    from betse_b200 import network as netlib
    from betse_b200.engine import TissueEngine
    cap = util.load_golden("mammal_ecm_net2")
    assert util.network_handlers(cap, kind) == [0, 1]
    descs = util.networks_of(cap, kind)
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
    specs = util.channels_of(cap, kind)
    for h, desc in enumerate(descs):
        comp = netlib.compile_network(desc, eng.Co, eng.M)
        eng.set_network(comp, handler=h)
        if h == 0:
            for c in specs:
                c["handler"] = 0
                c["mod_prog"] = comp["mod_index"][comp["chan_names"].index(c["name"])]
    eng.set_channels(specs, phase_init=(kind == "init"))
    n = 0
    snaps = util.snap_steps(cap, kind)
    for K in snaps:
        last = K == snaps[-1]
        while n < K:
            util.apply_schedule(eng, cap, kind, n + 1)
            assert not (eng.step(1, diag=(last and n + 1 == K)) & (3 | 16))
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        # Jmem / Jn: each MasterOfNetworks publishes its OWN extra_J_mem, the gene network's (no channels: zero) last — the
        # channels' currents of the general network are not in sim.extra_J_mem (networks.py:2971-2977)
        extra = ["Jmem", "Jn", "I_mem"] if last else []
        got = eng.download([f for f in list(util.STATE) + util.ENV_STATE + extra if f in ref])
        tols = util.gpu_tolerances(cap, kind, ref)
        for f, a in got.items():
            err = float(np.max(np.abs(np.asarray(a).reshape(np.shape(ref[f])) - ref[f])))
            assert err <= tols[f], (kind, K, f, err, tols[f])
        for h, desc in enumerate(descs):
            c = eng.network_state(h)
            for k, nme in enumerate(desc["species"]):
                assert util.rel_err(c[k], ref["net%d.c_cells" % h][k]) <= 1e-10, (kind, K, h, nme)
    eng.close()


def _retarget(desc, C, M, E, rng, targets_every=7):
    from betse_b200 import synth
    return synth.retarget_network(desc, C, M, E, rng, targets_every)


@pytest.mark.parametrize("fixture", ["mammal_ecm_net_trans", "mammal_ecm_net_pump", "mammal_ecm_net_lig", "mammal_ecm_net_envq",
                                     "mammal_ecm_net_intra"])
def test_network_features_vs_oracle_synthetic_30k(fixture):
    """The membrane / extracellular / transporter / pump / ligand-gate kernels at a size and grid shape the fixtures do
    not have (30 k cells, 174 x 174 grid): the recorded rate laws re-targeted to a synthetic tissue, 8 steps vs the oracle."""
    from betse_b200 import network as netlib
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    cap = util.load_golden(fixture)
    mesh, p, st = synth.make_tissue(30_000)
    p["substances_affect_charge"] = int(cap["sim.p.substances_affect_charge"])
    C, M = len(mesh["cell_vol"]), len(mesh["mem_sa"])
    E = int(np.prod(mesh["grid_shape"]))
    desc = _retarget(util.networks_of(cap, "sim")[0], C, M, E, np.random.default_rng(5))
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    ora = OracleSim(mesh, p, st, networks=[desc])
    ora.diagnostics = False
    ora.update_V()
    eng.set_network(netlib.compile_network(desc, eng.Co, eng.M), handler=0)
    eng.set_channels([])
    for n in range(8):
        assert not (eng.step(1) & (3 | 16))
        ora.step()
    got = eng.download(["cc_cells", "cc_env", "vm", "gjopen"])
    for f, a in got.items():
        r = np.asarray(getattr(ora, f))
        scale = max(float(np.max(np.abs(r))), 1e-300)
        tol = 1e-10 * scale if f != "vm" else max(1e-10 * scale, 4e-12)
        assert float(np.max(np.abs(a.reshape(r.shape) - r))) <= tol, (f, float(np.max(np.abs(a.reshape(r.shape) - r))), tol)
    c, ce = eng.network_state(0), eng.network_env_state(0)
    net = ora.networks[0]
    for k, nme in enumerate(desc["species"]):
        assert util.rel_err(c[k], net.c[nme]) <= 1e-10, nme
        if nme in net.c_env:
            assert util.rel_err(ce[k], net.c_env[nme]) <= 1e-10, (nme, "env")
    if np.any(desc.get("intra_on", 0)):
        cm = eng.network_mem_state(0)
        for k, nme in enumerate(desc["species"]):
            assert util.rel_err(cm[k], net.cmem[nme]) <= 1e-10, (nme, "mems")
    eng.close()
