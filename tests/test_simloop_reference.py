"""The drop-in seam against the REAL reference objects (build container only): `install()` rebinds
Simulator._run_sim_core_loop (betse/science/sim.py:1064-1075) and the shim builds the engine inputs
from a live Simulator / Cells / Parameters — checked up to the point where the CUDA engine would be
created (no GPU here), with a recording stand-in for TissueEngine."""
import numpy as np
import pytest

pytestmark = pytest.mark.reference


class _Stop(Exception):
    pass


def test_install_and_engine_inputs_from_live_reference(monkeypatch, tmp_path):
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse.science.sim import Simulator
    from betse_b200 import simloop
    from tests.golden.make_golden import CHANNELS, SMALL, _m

    seen = {}

    class FakeEngine:
        def __init__(self, mesh, params, state, device=0, partition=None):
            seen.update(mesh=mesh, params=params, state=state)
            self.ions = [str(x) for x in params["ions"]]

        def set_channels(self, specs, phase_init=False, affect_charge=None):
            seen.update(specs=specs, phase_init=phase_init, affect_charge=affect_charge)
            raise _Stop()

    monkeypatch.setattr(simloop, "TissueEngine", FakeEngine)
    orig = Simulator._run_sim_core_loop
    simloop.install()
    try:
        assert Simulator._run_sim_core_loop is not orig
        mods = _m(SMALL, {"cutting event": {"event happens": False}, "general options": {"ion profile": "mammal"},
                          "general network": {"implement network": True, "biomolecules": [], "channels": CHANNELS}})
        fn = refrun.write_config(str(tmp_path), mods)
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        with pytest.raises(_Stop):
            runner.init()
    finally:
        simloop.uninstall()
    assert Simulator._run_sim_core_loop is orig
    mesh, prm, st = seen["mesh"], seen["params"], seen["state"]
    M, C = len(mesh["mem_sa"]), len(mesh["cell_vol"])
    assert mesh["cell_mem_ptr"][-1] == M and len(mesh["nn_i"]) == M and len(mesh["map_mem2ecm"]) == M
    assert [str(x) for x in prm["ions"]] == ["Na", "K", "Cl", "Ca", "P", "M"]
    assert st["cc_cells"].shape == (6, C) and st["cc_at_mem"].shape == (6, M) and st["vm"].shape == (M,)
    assert st["cc_env"].shape[0] == 6 and st["Dm_cells"].shape == (6, M)
    assert seen["phase_init"] is True and seen["affect_charge"] is True
    assert [c["model"] for c in seen["specs"]] == ["Nav1p3", "Kv1p5", "KLeak", "Cav1p2"]
    assert [c["init_active"] for c in seen["specs"]] == [False, False, True, False]
    assert all(len(c["m"]) == M for c in seen["specs"])


def test_network_inputs_from_live_reference(monkeypatch, tmp_path):
    """General network with substances: the shim compiles the live MasterOfNetworks (strings evaluated against the
    real objects) and hands substance-modulated channels their program indices."""
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import simloop
    from tests.golden import make_golden as mg

    seen = {}

    class FakeEngine:
        def __init__(self, mesh, params, state, device=0, partition=None):
            self.ions = [str(x) for x in params["ions"]]
            self.Co, self.M = len(mesh["cell_vol"]), len(mesh["mem_sa"])

        def set_network(self, comp, handler=0):
            seen["net"] = (handler, comp)

        def set_channels(self, specs, phase_init=False, affect_charge=None):
            seen.update(specs=specs)
            raise _Stop()

    monkeypatch.setattr(simloop, "TissueEngine", FakeEngine)
    simloop.install()
    try:
        fn = refrun.write_config(str(tmp_path), mg.SCENARIOS["mammal_ecm_net"]["mods"])
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        with pytest.raises(_Stop):
            runner.init()
    finally:
        simloop.uninstall()
    handler, comp = seen["net"]
    assert handler == 0 and comp["species"] == ["G1", "G2", "G3", "X"]
    assert len(comp["rate_programs"]) == 5 and comp["mod_index"] == [-1, 5, 6, -1]
    assert list(comp["Dgj"]) == [-1.0, -1.0, -1.0, 1e-15]
    assert [c["mod_prog"] for c in seen["specs"]] == [-1, 5, 6, -1]
    assert comp["growth_mask"][3].sum() == 4 and comp["growth_mask"][:3].all()


def test_unsupported_network_is_refused(monkeypatch, tmp_path):
    """A substance with intracellular transport ('update intracellular') is outside the implemented subset: refused with the reason."""
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import simloop
    from betse_b200.capi import BetseB200Error
    from tests.golden import make_golden as mg
    import copy
    mods = copy.deepcopy(mg.SCENARIOS["mammal_ecm_net"]["mods"])
    mods["general network"]["biomolecules"][3]["update intracellular"] = True
    simloop.install()
    try:
        fn = refrun.write_config(str(tmp_path), mods)
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        with pytest.raises(BetseB200Error) as e:
            runner.init()
    finally:
        simloop.uninstall()
    assert "update intracellular" in str(e.value)


def _run_try(tmp_path, use_dropin, monkeypatch=None, mods=None, tweak_p=None):
    """`betse try` (seed + init + sim of the SHIPPED default config, cutting event included) -> the Simulator
    after the SIM phase.  ``use_dropin``: through betse_b200.simloop with the device replaced by the CPU oracle."""
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import network as netlib
    from betse_b200 import simloop
    from tests.oracle_engine import OracleEngine

    engines = []
    if use_dropin:
        real_compile = netlib.compile_network

        def compile_rec(desc, nc, nm, resolver=None):
            rec = {}

            def res(text):
                v = resolver(text)
                rec[text] = np.asarray(v, dtype=np.float64) if np.ndim(v) else float(v)
                return v
            comp = real_compile(desc, nc, nm, res)
            comp["_desc"] = dict(desc, static=rec)
            return comp

        class Eng(OracleEngine):
            def __init__(self, *a, **k):
                super().__init__(*a, **k)
                engines.append(self)

        monkeypatch.setattr(netlib, "compile_network", compile_rec)
        monkeypatch.setattr(simloop, "TissueEngine", Eng)
        simloop.install()
    try:
        fn = refrun.write_config(str(tmp_path), mods or {})
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        if tweak_p:
            tweak_p(p)
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        runner.init()
        phase = runner.sim()
    finally:
        if use_dropin:
            simloop.uninstall()
    return phase.sim, phase.cells, engines


@pytest.mark.parametrize("variant", ["shipped", "configs0_noecm"])
def test_betse_try_through_the_dropin_matches_the_reference(variant, monkeypatch, tmp_path):
    """The whole host side of the drop-in against the reference's own run of its shipped default configuration:
    INIT (500 steps) and SIM (350 steps, cutting event at the first step, three voltage-gated channels, substance X),
    sampled-step storage (`vm_time`, `cc_time`, ...) as `write2storage` leaves it.  The engine here is the CPU
    oracle (tests/oracle_engine.py); the CUDA engine is held to the same recorded run in tests/test_full_run.py."""
    # "configs0_noecm": BASELINE configs[0] to the letter (Na/K/Cl/Ca/P(+M) profile, no extracellular spaces)
    mods = {} if variant == "shipped" else {"general options": {"ion profile": "mammal", "simulate extracellular spaces": False}}
    ref_sim, ref_cells, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, new_cells, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert len(engines) == 2                                  # one engine per phase
    assert len(new_cells.mem_i) == len(ref_cells.mem_i) < engines[0].M      # the cut happened, on both paths alike
    assert engines[1].M == len(ref_cells.mem_i)               # the SIM engine was built from the post-cut mesh
    assert len(new_sim.time) == len(ref_sim.time) and np.allclose(new_sim.time, ref_sim.time, rtol=0, atol=0)
    assert len(ref_sim.vm_time) >= 30
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-6                  # BASELINE.json: Vmem traces within 1e-6 V
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))
    for name in ("cc_time", "cc_env_time", "gjopen_time", "vm_ave_time", "rho_cells_time", "I_mem_time",
                 "venv_time", "efield_gj_x_time", "rate_NaKATP_time", "I_tot_x_time", "I_tot_y_time",
                 "I_cell_x_time"):
        got, want = getattr(new_sim, name), getattr(ref_sim, name)
        assert len(got) == len(want) and len(want) >= 30, name
        for a, r in zip(got, want):
            a, r = np.asarray(a, dtype=float), np.asarray(r, dtype=float)
            assert a.shape == r.shape, name
            assert np.max(np.abs(a - r)) <= 1e-7 * max(np.max(np.abs(r)), 1e-300), name
    # final state left on the Simulator
    for f in ("cc_cells", "cc_env", "vm", "gjopen"):
        a, r = np.asarray(getattr(new_sim, f)), np.asarray(getattr(ref_sim, f))
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r)), f
    x_new = new_sim.molecules.core.molecules["X"].c_cells
    x_ref = ref_sim.molecules.core.molecules["X"].c_cells
    assert np.max(np.abs(x_new - x_ref)) <= 1e-9 * np.max(np.abs(x_ref))


def test_voltage_event_through_the_dropin_matches_the_reference(monkeypatch, tmp_path):
    """The external-voltage event (tissue/event/tisevevolt.py): fire_events ramps sim.bound_V on the host, the loop
    forwards it (TissueEngine.set_bound_V -> Phi_b re-solved by the engine), Vmem follows the reference's."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_volt"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert "bound_V" in engines[1].sets
    assert len(new_sim.vm_time) == len(ref_sim.vm_time) >= 30
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))
    assert np.max(np.abs(ref_sim.vm_time[0] - ref_sim.vm_time[5])) > 1e-4       # plateau vs after the event: it acted
    for a, r in zip(new_sim.cc_time, ref_sim.cc_time):
        assert np.max(np.abs(np.asarray(a) - np.asarray(r))) <= 1e-9 * np.max(np.abs(np.asarray(r)))


def test_env_substances_through_the_dropin_match_the_reference(monkeypatch, tmp_path):
    """Substances that cross the membrane and move through the extracellular grid (Molecule.transport ->
    stb.molecule_mover): the shim derives D_env (D_env_weight, TJ factor), c_bound and the env concentrations from
    the live objects, and copies c_env back for write_data (networks.py:4210-4226)."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_net_env"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    for name in ("S1", "S2", "S3", "G1"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        assert len(a.c_cells_time) == len(r.c_cells_time) >= 30
        for x, y in zip(a.c_cells_time + a.c_env_time, r.c_cells_time + r.c_env_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))


def test_tight_junction_modulators_through_the_dropin_match_the_reference(monkeypatch, tmp_path):
    """run_loop_modulators with target 'TJ' (networks.py:3301-3317): the shim hands over sim.TJ_targets, sim.D_env and the
    modulators' ion; the ion concentrations outside the cells (which the modulated barrier shapes), the substances and the
    sim.TJ_modulator the phase ends with — what the NEXT phase starts from — against the reference's own run."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_net_tj"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    for name in ("S1", "S2", "S3"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        assert len(a.c_env_time) == len(r.c_env_time) >= 30
        for x, y in zip(a.c_env_time, r.c_env_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    for a, r in zip(new_sim.cc_env_time, ref_sim.cc_env_time):
        assert np.max(np.abs(np.asarray(a) - np.asarray(r))) <= 1e-9 * np.max(np.abs(np.asarray(r)))
    tj_new, tj_ref = np.asarray(new_sim.TJ_modulator), np.asarray(ref_sim.TJ_modulator)
    assert float(np.ptp(tj_ref)) > 0.5
    assert np.max(np.abs(tj_new - tj_ref)) <= 1e-9 * np.max(np.abs(tj_ref))


def test_extracellular_reaction_through_the_dropin_matches_the_reference(monkeypatch, tmp_path):
    """A reaction outside the cells (write_reactions_env, networks.py:1830-2088; applied at the top of run_loop,
    networks.py:2872-2889): the shim compiles reaction_eval_string as an extracellular-zone rate law and hands over the
    substance rows of reaction_matrix_env; the env concentrations of reactant and product against the reference's own run."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_net_envrx"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    for name in ("S1", "S2", "S3"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        assert len(a.c_env_time) == len(r.c_env_time) >= 30
        for x, y in zip(a.c_env_time + a.c_cells_time, r.c_env_time + r.c_cells_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    s2 = ref_sim.molecules.core.molecules["S2"].c_env_time
    assert float(np.max(s2[-1]) - np.max(s2[0])) > 1e-4            # the reaction did produce S2 out there


def test_gene_network_through_the_dropin_matches_the_reference(monkeypatch, tmp_path):
    """BASELINE configs[3]: the shipped gene regulatory network (extra_configs/grn_basic.yaml) is the SECOND handler
    (sim.grn.core); the shim compiles it from the live MasterOfGenes and writes the genes back for write_data."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_grn"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert list(engines[1].networks) == [1]
    for name in ("Gene 1", "Gene 2", "Gene 3"):
        a, r = new_sim.grn.core.molecules[name], ref_sim.grn.core.molecules[name]
        assert len(a.c_cells_time) == len(r.c_cells_time) >= 30
        for x, y in zip(a.c_cells_time + a.c_mems_time, r.c_cells_time + r.c_mems_time):   # c_mems: 'update intracellular' is on
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * np.max(np.abs(np.asarray(y))), name
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))


def test_instability_raises_the_reference_exception(monkeypatch, tmp_path):
    """NaN reported by the engine's status word -> BetseSimUnstableException out of the drop-in, caught by the
    reference's own run_sim_core, which pickles the partial phase and re-raises (sim.py:1104-1128)."""
    import os
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.exceptions import BetseSimUnstableException
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import capi, network as netlib, simloop
    from tests.golden.make_golden import NO_NET, SMALL, _m
    from tests.oracle_engine import OracleEngine

    class Eng(OracleEngine):
        def step(self, n=1, diag=False):
            st = super().step(n, diag)
            self._steps = getattr(self, "_steps", 0) + n
            if self._steps >= 35:                               # after a few sampled steps of the INIT phase (100 steps, every 10th)
                self._sim().vm[3] = np.nan
                return capi.STATUS_NAN_VM
            return st

    monkeypatch.setattr(simloop, "TissueEngine", Eng)
    simloop.install()
    try:
        fn = refrun.write_config(str(tmp_path), _m(NO_NET, SMALL))
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        with pytest.raises(BetseSimUnstableException):
            runner.init()
    finally:
        simloop.uninstall()
    assert os.path.exists(p.init_pickle_filename)             # the partial phase was saved before the re-raise


def test_substance_events_through_the_dropin_match_the_reference(monkeypatch, tmp_path):
    """Molecule.update_boundary / cell_clamp_method (networks.py:2929-2933): the loop evaluates the schedules of every
    step's time on the host and hands the values to the engine; both phases."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_net_events"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert "net_events" in engines[0].sets and "net_events" in engines[1].sets
    for name in ("B1", "C1"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        assert len(a.c_cells_time) == len(r.c_cells_time) >= 30
        for x, y in zip(a.c_cells_time + a.c_env_time, r.c_cells_time + r.c_env_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    assert abs(new_sim.molecules.core.molecules["B1"].c_bound - ref_sim.molecules.core.molecules["B1"].c_bound) <= 1e-12


def test_dynamic_noise_through_the_dropin_matches_the_reference(monkeypatch, tmp_path):
    """Dynamic noise (sim.py:1322-1339): the loop draws protein_noise_flux from NumPy's global stream every SIM step, as
    the reference does, and hands it to the engine (TissueEngine.set_noise_flux); with the same seed both runs walk the
    protein concentration alike."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_dynnoise"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert engines[1].sets.count("noise") >= 100               # one draw per SIM step
    assert "noise" not in engines[0].sets                      # INIT draws nothing
    iP = ref_sim.iP
    assert np.ptp(ref_sim.cc_time[-1][iP]) > 0                 # the walk happened
    assert np.array_equal(new_sim.protein_noise_flux, ref_sim.protein_noise_flux)
    for a, r in zip(new_sim.cc_time, ref_sim.cc_time):
        a, r = np.asarray(a), np.asarray(r)
        assert np.max(np.abs(a[iP] - r[iP])) <= 1e-10 * np.max(np.abs(r[iP]))
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))


def test_intracellular_transport_through_the_dropin_matches_the_reference(monkeypatch, tmp_path):
    """'update intracellular' on for charged, membrane- and gap-junction permeable substances: the membrane values
    (Molecule.cc_at_mem -> c_mems_time) are transported state that the shim uploads, the engine advances and write_data
    stores (networks.py:5714-5806; sim_toolbox.py:962-1005, 1183-1185)."""
    from tests.golden import make_golden as mg
    mods = mg.SCENARIOS["mammal_ecm_net_intra"]["mods"]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    for name in ("A", "B", "Cp", "D"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        assert len(a.c_cells_time) == len(r.c_cells_time) >= 30
        for x, y in zip(a.c_cells_time + a.c_mems_time, r.c_cells_time + r.c_mems_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-9 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))


_GLOBAL_BLOCKS = {"block NaKATP pump": {"event happens": True, "change start": 2.0e-3, "change finish": 2.0e-2, "change rate": 1.0e-3},
                  "block gap junctions": {"event happens": True, "change start": 1.0e-3, "change finish": 1.5e-2,
                                          "change rate": 1.0e-3, "random fraction": 60}}
_BATH = {"change K env": {"event happens": True, "change start": 2.0e-3, "change finish": 2.0e-2, "change rate": 1.0e-3, "multiplier": 5},
         "change Na env": {"event happens": True, "change start": 5.0e-3, "change finish": 2.5e-2, "change rate": 2.0e-3, "multiplier": 2}}


def test_global_block_events_through_the_dropin_match_the_reference(monkeypatch, tmp_path):
    """Global scheduled interventions (tishandler.py:779-790).  'block NaKATP pump' REBINDS sim.NaKATP_block — an array
    np.ones(mdl) at loop entry (sim.py:848-849) — to a scalar every step, 'block gap junctions' writes a membrane subset
    of sim.gj_block in place.  Both must reach the engine (round-1 advisor finding: the scalar used to lose against the
    stale device array)."""
    from tests.golden.make_golden import NO_NET, SMALL, _m
    mods = _m(NO_NET, SMALL, _GLOBAL_BLOCKS, {"general options": {"ion profile": "mammal"}})
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=mods)
    assert "NaKATP_block" in engines[1].sets and "gj_block" in engines[1].sets
    # the pump block acted: its rate fell against the first sample
    assert np.max(np.abs(ref_sim.rate_NaKATP_time[10])) < 0.5 * np.max(np.abs(ref_sim.rate_NaKATP_time[0]))
    assert len(new_sim.vm_time) == len(ref_sim.vm_time) >= 30
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-9 * np.max(np.abs(r))
    for name in ("cc_time", "cc_env_time", "gjopen_time", "rate_NaKATP_time"):
        for a, r in zip(getattr(new_sim, name), getattr(ref_sim, name)):
            a, r = np.asarray(a, dtype=float), np.asarray(r, dtype=float)
            assert np.max(np.abs(a - r)) <= 1e-8 * max(np.max(np.abs(r)), 1e-300), name


def test_bath_events_without_ecm_fail_like_the_reference(monkeypatch, tmp_path):
    """'change K/Na/Cl env' without extracellular spaces is dead code in the reference: tishandler.py:761/767/777 read
    p.conc_env_k / conc_env_cl / conc_env_na, which parameters.py only declares (306-312) and never sets, so the first
    SIM step raises AttributeError.  The drop-in runs the reference's own fire_events and must fail the same way, not
    compute something else (round-1 advisor finding; TissueEngine.set_bath is what would serve the event if it ran)."""
    from tests.golden.make_golden import NO_NET, SMALL, _m
    mods = _m(NO_NET, SMALL, _BATH, {"general options": {"ion profile": "mammal", "simulate extracellular spaces": False}})
    with pytest.raises(AttributeError, match="conc_env_k"):
        _run_try(tmp_path / "ref", False, mods=mods)
    (tmp_path / "new").mkdir()
    with pytest.raises(AttributeError, match="conc_env_k"):
        _run_try(tmp_path / "new", True, monkeypatch, mods=mods)


@pytest.mark.parametrize("scenario", ["fast_basic", "fast_chan"])       # fast_chan: run_fast_loop_channels (networks.py:3217-3280)
def test_fast_solver_through_the_dropin_matches_the_reference(monkeypatch, tmp_path, scenario):
    """`solver options: type: fast` (sim.py:1068-1070): install() rebinds Simulator._run_fast_sim_core_loop as well; both
    phases of the reference's own run against the drop-in's host logic (events, sampling, the time series the fast loop
    appends itself, sim.py:1597-1628) over the oracle."""
    from tests.golden import make_golden as mg
    sc = mg.SCENARIOS[scenario]
    ref_sim, _, _ = _run_try(tmp_path / "ref", False, mods=sc["mods"], tweak_p=sc["tweak_p"])
    (tmp_path / "new").mkdir()
    new_sim, _, engines = _run_try(tmp_path / "new", True, monkeypatch, mods=sc["mods"], tweak_p=sc["tweak_p"])
    assert len(engines) == 2
    for name in ("vm_time", "vm_ave_time", "gjopen_time", "I_cell_x_time", "I_cell_y_time", "efield_gj_x_time",
                 "efield_gj_y_time", "time"):
        got, want = getattr(new_sim, name), getattr(ref_sim, name)
        assert len(got) == len(want) >= 10, name
        for a, r in zip(got, want):
            a, r = np.asarray(a, dtype=float), np.asarray(r, dtype=float)
            assert a.shape == r.shape, name
            assert np.max(np.abs(a - r)) <= 1e-9 * max(np.max(np.abs(r)), 1e-300), name
    for f in ("vm", "vm_ave", "gjopen", "Emx", "Emy", "Jn"):
        a, r = np.asarray(getattr(new_sim, f)), np.asarray(getattr(ref_sim, f))
        assert np.max(np.abs(a - r)) <= 1e-9 * max(np.max(np.abs(r)), 1e-300), f
    if scenario == "fast_chan":
        # what MasterOfNetworks.write_data stored for the channels at the sampled steps (networks.py:4244-4250)
        for name, ch in ref_sim.molecules.core.channels.items():
            got, want = new_sim.molecules.core.channels[name].flux_time, ch.flux_time
            assert len(got) == len(want) >= 10, name
            scale = max(float(np.max(np.abs(w))) for w in want)
            for a, r in zip(got, want):
                assert np.max(np.abs(np.asarray(a) - np.asarray(r))) <= 1e-9 * max(scale, 1e-300), name
