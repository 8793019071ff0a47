"""GPU vs oracle on synthetic tissues at BASELINE.json sizes (the oracle itself is pinned to the
real reference by tests/test_oracle_golden.py), plus size-independent properties at 1 M cells."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu


def _pair(n_cells, profile="mammal", ecm=True, ragged=0.0, **over):
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    mesh, p, st = synth.make_tissue(n_cells, profile=profile, ecm=ecm, overrides=over or None, ragged=ragged)
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    ora = OracleSim(mesh, p, st)
    ora.diagnostics = False
    ora.update_V()
    return mesh, p, eng, ora


def _compare(eng, ora, mesh, p, ecm, tag):
    fields = ["cc_cells", "cc_at_mem", "gjopen", "vm", "rho_cells", "cc_env"]
    if ecm:
        fields += ["E_env_x", "E_env_y", "v_env", "rho_env"]
    got = eng.download(fields)
    ref = {f: np.asarray(getattr(ora, f)) for f in fields}
    cap = {"x.p." + k: np.asarray(v) for k, v in p.items()}
    cap.update({"cells." + k: np.asarray(v) for k, v in mesh.items()})
    cap["x.s0.zs"] = ora.zs
    cap["x.s0.D_gj"] = ora.D_gj
    tol = util.gpu_tolerances(cap, "x", ref)
    for f in fields:
        err = float(np.max(np.abs(got[f].reshape(ref[f].shape) - ref[f])))
        assert err <= tol[f], (tag, f, err, tol[f])


@pytest.mark.parametrize("n_cells,profile,ecm,steps", [
    (10_000, "mammal", True, 20),      # BASELINE configs[1]
    (10_000, "basic", True, 5),
    (10_000, "mammal", False, 10),
    (100_000, "mammal", True, 3),      # BASELINE configs[2] size (ion path)
    (1, "mammal", True, 10),           # edge: the smallest tissue (4 cells, one partial warp tile, 7x7 grid: every env tile is an edge tile)
    (1, "basic", False, 10),
    (30, "mammal", True, 10),          # edge: fewer cells than one CTA's worth, most membranes on the cluster boundary
])
def test_gpu_vs_oracle(n_cells, profile, ecm, steps):
    mesh, p, eng, ora = _pair(n_cells, profile, ecm)
    _compare(eng, ora, mesh, p, ecm, "entry")
    for n in range(steps):
        st = eng.step(1)
        ora.step()
        assert not (st & 3)
    _compare(eng, ora, mesh, p, ecm, "step%d" % steps)
    eng.close()


@pytest.mark.parametrize("n_cells,steps", [(10_000, 10), (100_000, 3)])
def test_ragged_sheet_vs_oracle(n_cells, steps):
    """3-6 membranes per cell (synth.drop_membrane_pairs): blocks of the cell pack with different heights and padded lanes."""
    mesh, p, eng, ora = _pair(n_cells, ragged=0.2)
    assert sorted(set(np.diff(mesh["cell_mem_ptr"]).tolist())) == [3, 4, 5, 6]
    for n in range(steps):
        assert not (eng.step(1) & 3)
        ora.step()
    _compare(eng, ora, mesh, p, True, "ragged step%d" % steps)
    eng.close()


def test_batched_steps_equal_single_steps():
    """betse_step(n) (CUDA-graph replay) == n x betse_step(1): bit-identical state."""
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    mesh, p, st = synth.make_tissue(20_000)
    out = []
    for batched in (False, True):
        eng = TissueEngine(mesh, p, st)
        eng.update_V()
        if batched:
            eng.step(12)
        else:
            for _ in range(12):
                eng.step(1)
        out.append(eng.download(["cc_cells", "cc_env", "vm", "gjopen", "E_env_x"]))
        eng.close()
    for f in out[0]:
        assert np.array_equal(out[0][f], out[1][f]), f


def test_full_size_one_step_and_properties():
    """BASELINE configs[4] size (1 M cells, 6 M membranes, ~1000^2 grid): one step against the
    oracle, then size-independent properties over more steps: run-to-run determinism and
    gap-junction mass conservation (every GJ flux has an equal and opposite partner)."""
    mesh, p, eng, ora = _pair(1_000_000)
    st = eng.step(1)
    ora.step()
    assert not (st & 3)
    _compare(eng, ora, mesh, p, True, "1M step1")
    eng.step(9, diag=True)
    a = eng.download(["cc_cells", "vm", "fluxes_gj"])
    eng.close()
    from betse_b200.engine import TissueEngine
    from betse_b200 import synth
    mesh2, p2, st2 = synth.make_tissue(1_000_000)
    eng2 = TissueEngine(mesh2, p2, st2)
    eng2.update_V()
    eng2.step(10, diag=True)
    b = eng2.download(["cc_cells", "vm", "fluxes_gj"])
    eng2.close()
    assert np.array_equal(a["cc_cells"], b["cc_cells"]) and np.array_equal(a["vm"], b["vm"])
    # antisymmetry of junctional flux: f_gj[m]*sa[m] == -f_gj[nn[m]]*sa[nn[m]]
    nn = np.asarray(mesh["nn_i"])
    sa = np.asarray(mesh["mem_sa"])
    fg = a["fluxes_gj"] * sa
    resid = np.max(np.abs(fg + fg[:, nn]))
    # each side evaluates D*g/L*(c_nb*A - c_own*B) with its own rounding: the two agree to a few
    # ulp of the (cancelling) products, i.e. eps*(D*g/L)*c*sa, not of the tiny net flux
    scale = (np.max(st2["D_gj"]) * p2["gj_surface"] / float(mesh2["gj_len"])) * np.max(a["cc_cells"]) * np.max(sa)
    assert resid <= 1e-9 * np.max(np.abs(fg)) + 64 * util.EPS * scale, (resid, scale)


def test_dropin_loop_sampling_contract():
    """run_sim_core_loop (the Simulator._run_sim_core_loop drop-in, sim.py:1132-1390) on a synthetic tissue:
    what write2storage sees at every sampled step equals the oracle's state at that step; sampled-step views are
    engine-owned staging, the arrays left on the Simulator at the end are not."""
    import bench
    from betse_b200 import simloop, synth
    from oracle.betse_oracle import OracleSim
    mesh, p, state = synth.make_tissue(3000)
    sim, phase = bench.namespaces(mesh, p, state)
    n = 25
    ts = np.linspace(0, n * p["dt"], n)
    sampled = set(ts[4::5].tolist())
    ora = OracleSim(mesh, p, state)
    ora.diagnostics = False
    ora.update_V()
    seen = []

    def w2s(t, cells, pp):
        k = int(np.argmin(np.abs(ts - t))) + 1
        seen.append((k, {f: np.copy(getattr(sim, f)) for f in
                         ("cc_cells", "cc_env", "vm", "vm_ave", "gjopen", "rho_cells", "I_mem", "E_gj_x", "v_env",
                          "E_env_x", "rate_NaKATP", "J_cell_x")}, sim.vm_ave))
    sim.write2storage = w2s
    simloop.run_sim_core_loop(sim, phase, ts, sampled, None)
    assert [k for k, _, _ in seen] == [5, 10, 15, 20, 25]
    assert len({id(v) for _, _, v in seen}) == len(seen)          # vm_ave is appended uncopied by the reference
    done = 0
    for k, got, _ in seen:
        while done < k:
            ora.step()
            done += 1
        for f in ("cc_cells", "cc_env", "gjopen"):
            assert util.rel_err(got[f].reshape(np.shape(getattr(ora, f))), getattr(ora, f)) < 1e-10, (k, f)
        assert np.max(np.abs(got["vm"] - ora.vm)) < 1e-10 * np.max(np.abs(ora.vm)) + 4e-12
        assert got["E_env_x"].shape == ora.E_env_x.shape
    # after the loop: complete, engine-independent arrays (diagnostics included)
    for f in ("cc_cells", "vm", "fluxes_mem", "Jn", "E_gj_x", "cc_at_mem", "I_mem", "E_env_x"):
        a = getattr(sim, f)
        while isinstance(a, np.ndarray) and a.base is not None:
            a = a.base
        assert isinstance(a, np.ndarray) and a.flags.owndata, f      # not a view of page-locked staging
    assert util.rel_err(sim.cc_cells, ora.cc_cells) < 1e-10
    tol = 4 * (1e-10 * np.max(np.abs(ora.vm)) + 4e-12) / ora.gj_len
    assert np.max(np.abs(sim.E_gj_x - ora.E_gj_x)) <= tol


def test_helmholtz_hodge_vs_oracle_rectangular_grid():
    """csrc/hh.cu against the oracle's DST solve on a tissue whose env grid is not square and large enough that the
    sine-matrix products span several 64x64 tiles (the golden worlds are 20x19)."""
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    mesh, p, st = synth.make_tissue(20_000)
    rng = np.random.default_rng(11)        # a rough extracellular field, so that real currents flow in the env grid
    st = dict(st)
    st["cc_env"] = np.asarray(st["cc_env"]) * (1.0 + 0.02 * rng.standard_normal(np.shape(st["cc_env"])))
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    ora = OracleSim(mesh, p, st)
    ora.diagnostics = False
    ora.update_V()
    for n in range(3):
        s = eng.step(1, diag=(n == 2))
        ora.diagnostics = (n == 2)
        ora.step()
        assert not (s & 3)
    got = eng.download(["J_env_x", "J_env_y", "Jtx", "Jty", "B_field"])
    sJ = max(np.max(np.abs(ora.Jtx)), np.max(np.abs(ora.Jty)))
    for f in ("J_env_x", "J_env_y", "Jtx", "Jty"):
        assert np.max(np.abs(got[f].ravel() - np.asarray(getattr(ora, f)).ravel())) <= 1e-9 * sJ, f
    assert np.max(np.abs(got["B_field"] - ora.B_field)) <= 1e-7 * np.max(np.abs(ora.B_field)) + 1e-9 * sJ * p["mu"] * 1e-3
    eng.close()


def test_dropin_loop_raises_on_instability():
    """Error convention of the seam: NaN in Vmem / a concentration is not an error code but a status word;
    the drop-in copies the partial state back and raises the reference's BetseSimUnstableException (a stand-in class
    where the reference is not importable), like stb.check_v / stb.no_negs would (sim_toolbox.py:332-344, 475-503), so
    that run_sim_core can still pickle what there is (sim.py:1104-1128)."""
    import bench
    from betse_b200 import capi, simloop, synth
    from betse_b200.engine import TissueEngine
    mesh, p, state = synth.make_tissue(2000)
    # the status word itself: a NaN concentration surfaces within one step
    bad = dict(state)
    cc = np.array(state["cc_cells"], dtype=float, copy=True)
    cc[1, 17] = np.nan
    bad["cc_cells"] = cc
    eng = TissueEngine(mesh, p, bad)
    st = eng.step(1)
    assert st & (capi.STATUS_NAN_VM | capi.STATUS_NAN_CONC), st
    eng.close()
    # through the drop-in: exception after the copy-back, not a silent continue
    sim, phase = bench.namespaces(mesh, p, bad)
    n = 12
    ts = np.linspace(0, n * p["dt"], n)
    Unstable = simloop._unstable_exception()
    with pytest.raises(Unstable):
        simloop.run_sim_core_loop(sim, phase, ts, set(ts[3::4].tolist()), None)
    assert np.isnan(np.asarray(sim.cc_cells)).any() or np.isnan(np.asarray(sim.vm)).any()   # the partial state came back
    assert sim.sampled == 0                                                                   # nothing was stored after it


def test_block_array_then_scalar_reaches_the_device():
    """'block NaKATP pump' (tishandler.py:789-790): sim.NaKATP_block is np.ones(mdl) at loop entry (sim.py:848-849) and
    is REBOUND to a scalar by fire_events; the device must follow the scalar, not keep the stale array of ones.  Same
    for sim.gj_block (array -> array subset rewritten in place)."""
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from oracle.betse_oracle import OracleSim
    mesh, p, st = synth.make_tissue(3000)
    M = len(mesh["mem_sa"])
    st = dict(st, NaKATP_block=np.ones(M), gj_block=np.ones(M))
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    ora = OracleSim(mesh, p, st)
    ora.diagnostics = False
    ora.update_V()
    gjb = np.ones(M)
    gjb[::3] = 0.25
    for n, (blk, gj) in enumerate([(1.0, None), (0.4, gjb), (0.0, gjb), (0.7, np.ones(M))]):
        eng.set_field("NaKATP_block", blk)                # scalar, like the reference's rebinding
        ora.NaKATP_block = np.asarray(blk, dtype=float)
        if gj is not None:
            eng.set_field("gj_block", gj)
            ora.gj_block = np.array(gj)
        for _ in range(3):
            assert not (eng.step(1) & 3)
            ora.step()
        _compare(eng, ora, mesh, p, True, "block phase %d" % n)
    eng.close()


def test_cell_kernel_equals_generic_kernel_bitwise(monkeypatch):
    """k_cell (lane per cell, cell pack) and the run-time-configured k_mem share their arithmetic (csrc/kmath.cuh):
    the states they produce must be bit-identical, on a uniform sheet and on a ragged reference mesh (3-7 membranes per
    cell: padded rows of the cell pack)."""
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from tests import util
from betse_b200 import synth
from betse_b200.engine import TissueEngine
out = {}
mesh, p, st = synth.make_tissue(5000)
eng = TissueEngine(mesh, p, st); eng.update_V(); eng.step(7)
out["synth"] = eng.download(["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "E_env_x"]); eng.close()
cap = util.load_golden("mammal_ecm")
mesh, p, s0 = util.group(cap, "cells."), util.group(cap, "sim.p."), util.group(cap, "sim.s0.")
eng = TissueEngine(mesh, p, s0); eng.step(7)
out["golden"] = eng.download(["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "E_env_x"]); eng.close()
np.savez(sys.argv[1], **{k + "." + f: a for k, d in out.items() for f, a in d.items()})
''' % util.ROOT
    import os
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, env in (("cell", {"BETSE_KCELL": "1"}), ("generic", {"BETSE_KCELL": "0", "BETSE_KMEM_GENERIC": "1"})):
            fn = os.path.join(d, tag + ".npz")
            subprocess.run([sys.executable, "-c", code, fn], check=True, env=dict(os.environ, **env))
            with np.load(fn) as z:
                res[tag] = {k: z[k] for k in z.files}
    assert res["cell"].keys() == res["generic"].keys() and len(res["cell"]) == 12
    for k in res["cell"]:
        assert np.array_equal(res["cell"][k], res["generic"][k]), k


def test_patch_kernel_equals_plain_cell_kernel_bitwise():
    """k_cell_patch (BETSE_PATCH=1: blocks composed from a spatial sort, membrane -> env exchange on chip for the env squares
    a patch owns) sums every env square in the order of the global membrane index like k_envacc_ell: the states must be
    bit-identical to the plain k_cell path, on the uniform sheet and on a ragged one."""
    import os
    import subprocess
    import sys
    import tempfile
    code = r'''
import sys, numpy as np
sys.path.insert(0, %r)
from betse_b200 import synth
from betse_b200.engine import TissueEngine
out = {}
for tag, kw in (("sheet", {}), ("ragged", {"ragged": 0.2})):
    mesh, p, st = synth.make_tissue(30000, **kw)
    eng = TissueEngine(mesh, p, st); eng.update_V(); eng.step(9)
    out[tag] = eng.download(["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "E_env_x", "v_env"]); eng.close()
np.savez(sys.argv[1], **{k + "." + f: a for k, d in out.items() for f, a in d.items()})
''' % util.ROOT
    res = {}
    with tempfile.TemporaryDirectory() as d:
        for tag, env in (("patch", {"BETSE_PATCH": "1"}), ("plain", {"BETSE_PATCH": "0"})):
            fn = os.path.join(d, tag + ".npz")
            subprocess.run([sys.executable, "-c", code, fn], check=True, env=dict(os.environ, **env))
            with np.load(fn) as z:
                res[tag] = {k: z[k] for k in z.files}
    assert res["patch"].keys() == res["plain"].keys() and len(res["patch"]) == 14
    for k in res["patch"]:
        assert np.array_equal(res["patch"][k], res["plain"][k]), k


def test_uniform_permeability_upload_equals_full_upload(monkeypatch):
    """betse_upload_state sends sim.Dm_cells as I scalars + a device fill when every row is one repeated value (capi.cu:
    rows_uniform): same device state as the full I*M upload (BETSE_DM_UNIFORM=0), bit for bit; a tissue with ONE membrane
    of different permeability takes the full path and differs from the uniform one."""
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    mesh, p, st = synth.make_tissue(20_000)
    out = []
    for mode in ("1", "0"):
        monkeypatch.setenv("BETSE_DM_UNIFORM", mode)
        eng = TissueEngine(mesh, p, st)
        eng.update_V()
        assert not (eng.step(12) & 3)
        out.append(eng.download(["cc_cells", "cc_env", "vm"]))
        eng.close()
    for f in out[0]:
        assert np.array_equal(out[0][f], out[1][f]), f
    monkeypatch.setenv("BETSE_DM_UNIFORM", "1")
    st2 = dict(st)
    st2["Dm_cells"] = np.array(st["Dm_cells"], copy=True)
    st2["Dm_cells"][1, -1] *= 50.0                       # the LAST membrane of the K row: found by the last scanning thread
    eng = TissueEngine(mesh, p, st2)
    eng.update_V()
    assert not (eng.step(12) & 3)
    got = eng.download(["cc_cells"])
    eng.close()
    assert not np.array_equal(got["cc_cells"], out[0]["cc_cells"])
