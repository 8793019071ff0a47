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
