"""Voltage-gated channels on the GPU (BASELINE configs[2]) through the C ABI:
against the REAL reference's recorded run of a general network with Nav1p3 / Kv1p5 / KLeak /
Cav1p2 (tests/golden/mammal_ecm_chan.npz), and against the oracle on a synthetic tissue."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind", ["init", "sim"])
@pytest.mark.parametrize("fixture", ["mammal_ecm_chan", "mammal_ecm_chan_multi", "mammal_ecm_chan_ml"])   # _multi: vg_funny HCN2 + cation leak (Na/K/Ca each); _ml: Morris-Lecar family
def test_channels_match_reference(fixture, kind):
    from betse_b200.engine import TissueEngine
    cap = util.load_golden(fixture)
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
    specs = util.channels_of(cap, kind)
    eng.set_channels(specs, phase_init=(kind == "init"))
    active = [c for c in specs if not (kind == "init" and not c["init_active"])]
    n = 0
    snaps = util.snap_steps(cap, kind)
    for K in snaps:
        last = K == snaps[-1]
        while n < K:
            assert not util.group(cap, "%s.sched.k%d." % (kind, n + 1))
            st = eng.step(1, diag=(last and n + 1 == K))
            assert not (st & 3)
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        fields = list(util.STATE) + util.ENV_STATE + (["Jmem", "fluxes_mem"] if last else [])
        got = eng.download([f for f in fields if f in ref])
        tols = util.gpu_tolerances(cap, kind, ref)
        for f, a in got.items():
            err = float(np.max(np.abs(np.asarray(a).reshape(np.shape(ref[f])) - ref[f])))
            assert err <= tols[f], (kind, K, f, err, tols[f])
        for k, c in enumerate(active):
            j = [s["name"] for s in specs].index(c["name"])
            stt = eng.channel_state(k)
            tg = c["targets"]
            for f in ("m", "h"):
                r = ref["chan%d.%s" % (j, f)]
                assert np.max(np.abs(stt[f][tg] - r)) <= 1e-10 * max(np.max(np.abs(r)), 1e-300), (kind, K, c["name"], f)
            r = ref["chan%d.P" % j]
            assert np.max(np.abs(stt["P"] - r)) <= 1e-10 * max(np.max(np.abs(r)), 1e-300), (kind, K, c["name"], "P")
            r = ref["chan%d.flux" % j]                      # chan_flux: the LAST conducted ion's flux (networks.py:3201)
            assert np.max(np.abs(stt["flux"] - r)) <= max(1e-10 * np.max(np.abs(r)), tols["fluxes_mem"] if "fluxes_mem" in tols else 0.0), (kind, K, c["name"], "flux")
    eng.close()


@pytest.mark.parametrize("n_cells,steps", [(10_000, 15), (100_000, 6)])        # 100 k cells: BASELINE configs[2] to the letter
def test_channels_vs_oracle_synthetic(n_cells, steps):
    """Synthetic tissue with the four channels of BASELINE configs[2] (mammal profile, vg Na / K / Ca + leak, Na/K pump)."""
    from betse_b200 import channels as chlib
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    mesh, p, st = synth.make_tissue(n_cells)
    p["substances_affect_charge"] = 1
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
    ora = OracleSim(mesh, p, st, channels=specs)
    ora.diagnostics = False
    ora.update_V()
    eng.set_channels(specs)
    for n in range(steps):
        s = eng.step(1)
        ora.step()
        assert not (s & 3)
    got = eng.download(["cc_cells", "cc_env", "vm", "gjopen"])
    for f, a in got.items():
        r = np.asarray(getattr(ora, f))
        scale = max(float(np.max(np.abs(r))), 1e-300)
        tol = 1e-10 * scale if f != "vm" else max(1e-10 * scale, 4e-12)
        assert float(np.max(np.abs(a.reshape(r.shape) - r))) <= tol, f
    for k, c in enumerate(ora.channels):
        stt = eng.channel_state(k)
        for f in ("m", "h", "P"):
            assert np.max(np.abs(stt[f] - c[f])) <= 1e-10 * max(np.max(np.abs(c[f])), 1e-300), (c["name"], f)
    eng.close()


def test_every_tabulated_model_on_the_device():
    """Each model of the gating table (betse_b200/channels.py, pinned to the reference's classes by
    tests/test_channels_table.py) through k_chan, 3 steps on a 2 k-cell tissue, against the table's NumPy evaluation in the
    oracle.  The synthetic tissue rests near 0 V, so every model also runs as copies shifted by -70.3, -24.7 and +15.4 mV
    (both sides of Cav3p1's tau cut-off, vg_ca.py:515-519)."""
    import copy
    from betse_b200 import channels as chlib
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    mesh, p, st = synth.make_tissue(2_000)
    ora0 = OracleSim(mesh, p, st)
    ora0.diagnostics = False
    ora0.update_V()
    base = sorted(chlib.MODELS)
    added = []
    try:
        for model in base:
            for dv in (-70.3, -24.7, 15.4):        # off the round voltages where the rates have removable 0/0 points
                name = "%s@%+.1f" % (model, dv)
                chlib.MODELS[name] = dict(copy.deepcopy(chlib.MODELS[model]), shift=chlib.MODELS[model]["shift"] + dv)
                added.append(name)
        n_run, bad = 0, []
        for model in base + added:
            ions, _ = chlib.ions_of(model)
            if any(i not in p["ions"] for i in ions):
                continue
            m0, h0 = chlib.initial_state(model, ora0.vm)
            spec = chlib.make_channel("c", model, 1.0e-16, m=m0, h=h0)
            ora = OracleSim(mesh, p, st, channels=[dict(spec)])
            ora.diagnostics = False
            ora.update_V()
            eng = TissueEngine(mesh, p, st)
            eng.update_V()
            eng.set_channels([spec])
            for n in range(3):
                assert not (eng.step(1) & 3), model
                ora.step()
            stt, c = eng.channel_state(0), ora.channels[0]
            for f in ("m", "h", "P"):
                err = np.max(np.abs(stt[f] - c[f])) / max(np.max(np.abs(c[f])), 1e-300)
                if err > 1e-10:
                    bad.append((model, f, float(err)))
            eng.close()
            n_run += 1
        assert not bad, bad
        assert n_run >= 4 * 40
    finally:
        for name in added:
            del chlib.MODELS[name]


def test_channel_cell_path_equals_one_pair_per_channel(monkeypatch):
    """k_chan_cell (one lane per cell, consecutive channels of different ions in one pass, channels.cu) is the same arithmetic in the
    same order as one k_chan / k_chan_env pair per channel: bit-identical state after 12 steps (BETSE_CHAN_CELL=0: pairs)."""
    from betse_b200 import channels as chlib
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    mesh, p, st = synth.make_tissue(20_000)
    p["substances_affect_charge"] = 1
    out = []
    for mode in ("1", "0"):
        monkeypatch.setenv("BETSE_CHAN_CELL", mode)
        eng = TissueEngine(mesh, p, st)
        eng.update_V()
        vm0 = eng.download(["vm"])["vm"]
        specs = []
        for name, model, dm in (("Nav", "Nav1p3", 2.0e-14), ("Kv", "Kv1p5", 1.0e-15), ("K_Leak", "KLeak", 0.6e-17),
                                ("Cav", "Cav1p2", 1.0e-15), ("HCN", "HCN2", 1.0e-16)):
            m0, h0 = chlib.initial_state(model, vm0)
            specs.append(chlib.make_channel(name, model, dm, m=m0, h=h0))
        eng.set_channels(specs)
        assert not (eng.step(12) & 3)
        got = eng.download(["cc_cells", "cc_env", "vm"])
        for k in range(len(specs)):
            got.update({"%s%d" % (f, k): a for f, a in eng.channel_state(k).items() if a is not None})
        out.append(got)
        eng.close()
    for f in out[0]:
        assert np.array_equal(out[0][f], out[1][f]), f
