"""The FAST (equivalent-circuit) solver on the GPU (Simulator._run_fast_sim_core_loop, sim.py:1454-1640)
through the C ABI: against the REAL reference's recorded runs (tests/golden/fast_*.npz) and against the oracle on a
synthetic tissue of BASELINE size."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

FIELDS = ["vm_ave", "gjopen", "vgj", "Jn", "J_cell_x", "J_cell_y", "E_cell_x", "E_cell_y", "Emx", "Emy"]


def _close(a, r, scale=None, tol=1e-10):
    s = max(float(np.max(np.abs(r))) if scale is None else scale, 1e-300)
    return float(np.max(np.abs(np.asarray(a) - np.asarray(r)))) <= tol * s


@pytest.mark.parametrize("kind", ["init", "sim"])
@pytest.mark.parametrize("fixture", util.GOLDEN_FAST)
def test_fast_solver_matches_reference(fixture, kind):
    from betse_b200.engine import TissueEngine
    cap = util.load_golden(fixture)
    s0 = util.group(cap, kind + ".s0.")
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), s0)
    specs = util.channels_of(cap, kind)                 # fast_chan: run_fast_loop_channels (networks.py:3217-3280)
    active = [c for c in specs if not (kind == "init" and not c["init_active"])]
    if specs:
        fc = util.fast_consts_of(cap, kind)
        eng.set_channels(specs, phase_init=(kind == "init"))
        eng.fast_set_channels(fc["cbar"], fc["rev_E"], fc["geo_conv"])
    eng.fast_setup(s0)
    n = 0
    for K in util.snap_steps(cap, kind):
        st = eng.fast_step(K - n, diag=True)
        assert not (st & 1)
        n = K
        ref = util.group(cap, "%s.k%d." % (kind, K))
        got = eng.fast_download(FIELDS)
        # the cell-centre currents are sums over a closed polygon (heavy cancellation): judged against their summands
        jscale = float(np.max(np.abs(ref["Jn"])))
        for f in FIELDS + ["vm"]:
            if f not in ref:
                continue
            scale = None
            if f.startswith("J_cell"):
                scale = jscale
            elif f.startswith("E_cell"):
                scale = jscale / (0.1 * float(np.min(s0["sigma_cell"])))
            assert _close(got[f], ref[f], scale), (kind, K, f, float(np.max(np.abs(got[f] - ref[f]))))
        for k, c in enumerate(active):
            j = [s["name"] for s in specs].index(c["name"])
            stt = eng.channel_state(k)
            for f in ("m", "h", "P", "flux"):
                r = ref.get("chan%d.%s" % (j, f))
                if r is not None:
                    a = stt[f][c["targets"]] if f in ("m", "h") else stt[f]
                    assert _close(a, r), (kind, K, c["name"], f)
    eng.close()


def test_fast_solver_vs_oracle_synthetic():
    """100 k cells: random leak circuit on the synthetic sheet, graph replay included (> 8 steps per call)."""
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleFastSim
    mesh, p, st = synth.make_tissue(100_000)
    C, M = len(mesh["cell_vol"]), len(mesh["mem_sa"])
    rng = np.random.default_rng(5)
    fs = {"vm_ave": rng.uniform(-0.07, 0.0, C), "gjopen": rng.uniform(0.2, 1.0, M), "G_Leak": rng.uniform(0.1, 1.2, C),
          "E_Leak": rng.uniform(-0.08, 0.0, C), "G_gj": 2.1 / mesh["num_mems"], "sigma_cell": rng.uniform(0.02, 0.03, C),
          "extra_J_mem": 1e-3 * rng.standard_normal(M), "gj_block": 1.0}
    eng = TissueEngine(mesh, p, st)
    eng.fast_setup(fs)
    ora = OracleFastSim(mesh, p, dict(st, **fs))
    for run in (1, 2, 19, 3):
        assert not (eng.fast_step(run, diag=True) & 1)
        for _ in range(run):
            ora.step()
        got = eng.fast_download(FIELDS)
        jscale = float(np.max(np.abs(ora.Jn)))
        for f in FIELDS:
            scale = jscale if f.startswith("J_cell") else (jscale / 0.002 if f.startswith("E_cell") else None)
            assert _close(got[f], getattr(ora, f), scale), (run, f)
    eng.close()
