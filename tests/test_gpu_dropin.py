"""The REAL drop-in on a GPU: `simloop.install()` rebinds Simulator._run_sim_core_loop (betse/science/sim.py:1064-1075) of
the unmodified reference, `SimRunner.seed/init/sim` (simrunner.py:93-296) runs the shipped default configuration
(`betse try`: 228 cells, basic ion profile, extracellular spaces, general network with substance X and three channels,
cutting event) once with the reference's own NumPy loop and once through the CUDA engine, and the stored time series are
compared.  Needs the reference tree: /root/reference (build container) or baseline/_ref (tools/install_reference.py
puts it there; it travels to the GPU box with gpurun) — skipped, visibly, when neither exists."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _have_reference():
    from oracle import refshim
    return refshim.reference_available()


def _run(tmp_path, use_dropin, mods=None, tweak_p=None):
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import simloop
    if use_dropin:
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
    return phase.sim, phase.cells


@pytest.mark.parametrize("variant", ["shipped", "mammal_ion_path"])
def test_betse_try_reference_loop_vs_cuda_dropin(variant, tmp_path):
    if not _have_reference():
        pytest.skip("reference tree absent: neither /root/reference nor baseline/_ref (run tools/install_reference.py)")
    from tests.golden.make_golden import NO_NET, SMALL, _m
    mods = {} if variant == "shipped" else _m(NO_NET, SMALL, {"general options": {"ion profile": "mammal"}})
    (tmp_path / "ref").mkdir()
    (tmp_path / "new").mkdir()
    ref_sim, ref_cells = _run(tmp_path / "ref", False, mods)
    new_sim, new_cells = _run(tmp_path / "new", True, mods)
    assert len(new_cells.mem_i) == len(ref_cells.mem_i)
    assert len(new_sim.time) == len(ref_sim.time) and len(ref_sim.vm_time) >= 30
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-6                       # BASELINE.json: final Vmem traces within 1e-6 V
        assert np.max(np.abs(a - r)) <= 1e-8 * np.max(np.abs(r))
    for name in ("cc_time", "cc_env_time", "gjopen_time", "vm_ave_time", "rho_cells_time", "I_mem_time", "venv_time"):
        got, want = getattr(new_sim, name), getattr(ref_sim, name)
        assert len(got) == len(want) >= 30, name
        for a, r in zip(got, want):
            a, r = np.asarray(a, dtype=float), np.asarray(r, dtype=float)
            assert a.shape == r.shape, name
            assert np.max(np.abs(a - r)) <= 1e-8 * max(np.max(np.abs(r)), 1e-300), name
    for f in ("cc_cells", "cc_env", "vm", "gjopen"):
        a, r = np.asarray(getattr(new_sim, f)), np.asarray(getattr(ref_sim, f))
        assert np.max(np.abs(a - r)) <= 1e-8 * np.max(np.abs(r)), f
    if variant == "shipped":
        x_new = new_sim.molecules.core.molecules["X"].c_cells
        x_ref = ref_sim.molecules.core.molecules["X"].c_cells
        assert np.max(np.abs(x_new - x_ref)) <= 1e-8 * np.max(np.abs(x_ref))


@pytest.mark.parametrize("scenario", ["fast_basic", "fast_chan"])       # fast_chan: run_fast_loop_channels (networks.py:3217-3280)
def test_fast_solver_reference_loop_vs_cuda_dropin(tmp_path, scenario):
    """`solver options: type: fast`: Simulator._run_fast_sim_core_loop of the unmodified reference against the drop-in
    that install() binds in its place (csrc/fast.cu), both phases, a K-leaky tissue profile driving gap-junction currents."""
    if not _have_reference():
        pytest.skip("reference tree absent: neither /root/reference nor baseline/_ref (run tools/install_reference.py)")
    from tests.golden import make_golden as mg
    sc = mg.SCENARIOS[scenario]
    (tmp_path / "ref").mkdir()
    (tmp_path / "new").mkdir()
    ref_sim, _ = _run(tmp_path / "ref", False, sc["mods"], sc["tweak_p"])
    new_sim, _ = _run(tmp_path / "new", True, sc["mods"], sc["tweak_p"])
    for name in ("vm_time", "vm_ave_time", "gjopen_time", "I_cell_x_time", "I_cell_y_time", "efield_gj_x_time", "time"):
        got, want = getattr(new_sim, name), getattr(ref_sim, name)
        assert len(got) == len(want) >= 10, name
        scale = max(float(np.max(np.abs(np.asarray(w, dtype=float)))) for w in want)
        if name.startswith(("I_cell", "efield")):
            # sums over a closed polygon: judged against the summands (the membrane currents), like tests/test_gpu_fast.py
            scale = max(scale, float(np.max(np.abs(ref_sim.Jn))) / (0.1 * float(np.min(ref_sim.sigma_cell)) if name.startswith("efield") else 1.0))
        for a, r in zip(got, want):
            assert np.max(np.abs(np.asarray(a, dtype=float) - np.asarray(r, dtype=float))) <= 1e-9 * max(scale, 1e-300), name


def test_env_zone_network_reference_loop_vs_cuda_dropin(tmp_path):
    """The extracellular zone of the networks on the GPU through the real drop-in: two tight-junction modulators
    (run_loop_modulators, target 'TJ', networks.py:3301-3317) and a reaction outside the cells (write_reactions_env,
    networks.py:1830-2088) in one general network — the env concentrations of ions and substances, Vmem and the
    sim.TJ_modulator the run ends with, against the unmodified reference's own loop."""
    if not _have_reference():
        pytest.skip("reference tree absent: neither /root/reference nor baseline/_ref (run tools/install_reference.py)")
    import copy
    from tests.golden import make_golden as mg
    mods = copy.deepcopy(mg.SCENARIOS["mammal_ecm_net_tj"]["mods"])
    mods["general network"]["reactions"] = copy.deepcopy(mg._ENV_RX)
    (tmp_path / "ref").mkdir()
    (tmp_path / "new").mkdir()
    ref_sim, _ = _run(tmp_path / "ref", False, mods)
    new_sim, _ = _run(tmp_path / "new", True, mods)
    assert len(new_sim.time) == len(ref_sim.time) and len(ref_sim.vm_time) >= 30
    for a, r in zip(new_sim.vm_time, ref_sim.vm_time):
        assert np.max(np.abs(a - r)) <= 1e-6
    for a, r in zip(new_sim.cc_env_time, ref_sim.cc_env_time):
        assert np.max(np.abs(np.asarray(a) - np.asarray(r))) <= 1e-8 * np.max(np.abs(np.asarray(r)))
    for name in ("S1", "S2", "S3"):
        a, r = new_sim.molecules.core.molecules[name], ref_sim.molecules.core.molecules[name]
        for x, y in zip(a.c_env_time, r.c_env_time):
            assert np.max(np.abs(np.asarray(x) - np.asarray(y))) <= 1e-8 * max(np.max(np.abs(np.asarray(y))), 1e-300), name
    tj_new, tj_ref = np.asarray(new_sim.TJ_modulator), np.asarray(ref_sim.TJ_modulator)
    assert float(np.ptp(tj_ref)) > 0.5 and np.max(np.abs(tj_new - tj_ref)) <= 1e-8 * np.max(np.abs(tj_ref))
