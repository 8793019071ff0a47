"""GPU parity against the REAL reference's recorded outputs (tests/golden/*.npz), through the
C-ABI.  Tolerance: BASELINE.json — per-step ion concentrations and Vmem within 1e-10 relative
(max-norm over the array); cancelling sums (charge, Vmem, currents) additionally get the
forward-error bound of the sum itself, see tests/util.py:gpu_tolerances."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-10


def _engine(cap, kind):
    from betse_b200.engine import TissueEngine
    return TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."),
                        util.group(cap, kind + ".s0."))


@pytest.mark.parametrize("name", [n for n in util.GOLDEN if "chan" not in n and "net" not in n and not n.startswith("default_try")])
@pytest.mark.parametrize("kind", ["init", "sim"])
def test_gpu_matches_reference(name, kind):
    cap = util.load_golden(name)
    eng = _engine(cap, kind)
    ecm = bool(int(cap[kind + ".p.is_ecm"]))
    snaps = util.snap_steps(cap, kind)
    n = 0
    worst = {}
    for K in snaps:
        last = K == snaps[-1]
        while n < K:
            util.apply_schedule(eng, cap, kind, n + 1)
            st = eng.step(1, diag=(last and n + 1 == K))
            assert not (st & 3), "instability flagged"
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        fields = list(util.STATE) + (util.ENV_STATE if ecm else [])
        if last:
            fields += [f for f in util.DIAG if f in ref and f not in ("rho_env_surf", "Eme")]
            if not ecm:      # no-ECM field diagnostics (ion_current.py:116-158): local field potential, its field, bath current
                fields = [f for f in fields if not f.startswith("fluxes_env")] + ["v_env", "E_env_x", "E_env_y"]
        got = eng.download([f for f in fields if f in ref])
        tols = util.gpu_tolerances(cap, kind, ref)
        for f, a in got.items():
            err = float(np.max(np.abs(np.asarray(a).reshape(np.shape(ref[f])) - ref[f])))
            rel = err / max(float(np.max(np.abs(ref[f]))), 1e-300)
            worst[f] = max(worst.get(f, 0.0), rel)
            assert err <= tols[f], (name, kind, K, f, "abs err %.3e > tol %.3e (rel %.2e)" % (err, tols[f], rel))
    print(name, kind, {k: float("%.1e" % v) for k, v in worst.items()})
    eng.close()
