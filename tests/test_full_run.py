"""BASELINE configs[0] as shipped (`betse try`): the unmodified default sim_config.yaml run through BOTH full phases
(500 INIT + 350 SIM timesteps; basic ion profile, ECM, general network with substance X and the Nav1p3 / Kv1p5 /
X-inhibited KLeak channels, the cutting event at the first SIM step) against the traces the REAL reference recorded
at every sampled step (tests/golden/default_try.npz, made by tests/golden/make_golden.py).

Bars (BASELINE.json): Vmem traces within 1e-6 V absolute at every sampled step of the whole run; concentrations are
additionally held to 1e-8 relative after hundreds of steps (the per-step 1e-10 bar is tests/test_gpu_golden.py's)."""
import numpy as np
import pytest

from tests import util

VM_ABS = 1.0e-6       # volts, BASELINE.json "final Vmem traces within 1e-6 V absolute"
CONC_REL = 1.0e-8


def _run_traces(obj, cap, kind, step, get):
    """Advance `obj` through the phase; return {field: [value at each traced step]}."""
    steps = [int(s) for s in cap[kind + ".trace.steps"]]
    n_total = int(max(util.snap_steps(cap, kind)))
    out = {"vm": [], "cc_cells": []}
    for n in range(1, n_total + 1):
        util.apply_schedule(obj, cap, kind, n)
        step(obj)
        if n in steps:
            for f in out:
                out[f].append(np.array(get(obj, f), dtype=float, copy=True))
    return out


def _check(cap, kind, got, who):
    for f, tol_abs in (("vm", VM_ABS), ("cc_cells", None)):
        ref = cap["%s.trace.%s" % (kind, f)]
        assert len(got[f]) == len(ref) and len(ref) >= 4
        for j, (a, r) in enumerate(zip(got[f], ref)):
            err = float(np.max(np.abs(a.reshape(r.shape) - r)))
            tol = tol_abs if tol_abs else CONC_REL * float(np.max(np.abs(r)))
            assert err <= tol, (who, kind, f, "sample %d" % j, err, tol)


FIXTURES = ["default_try", "default_try_noecm"]   # shipped default (ECM on); BASELINE configs[0] to the letter (mammal profile, no ECM)


@pytest.mark.parametrize("kind", ["init", "sim"])
@pytest.mark.parametrize("fixture", FIXTURES)
def test_oracle_full_run(fixture, kind):
    from oracle.betse_oracle import OracleSim
    cap = util.load_golden(fixture)
    o = OracleSim(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."),
                  channels=util.channels_of(cap, kind), phase_init=(kind == "init"), networks=util.networks_of(cap, kind),
                  net_handlers=util.network_handlers(cap, kind))
    o.diagnostics = False
    got = _run_traces(o, cap, kind, lambda x: x.step(), getattr)
    _check(cap, kind, got, "oracle")
    # the oracle itself is pinned far tighter than the bar it checks others against
    K = max(util.snap_steps(cap, kind))
    ref = util.group(cap, "%s.k%d." % (kind, K))
    for f in ("cc_cells", "cc_env", "vm", "gjopen"):
        assert util.rel_err(getattr(o, f), ref[f]) < 1e-9, (kind, f, util.rel_err(getattr(o, f), ref[f]))
    for k, nme in enumerate(o.networks[0].species):
        assert util.rel_err(o.networks[0].c[nme], ref["net0.c_cells"][k]) < 1e-9


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["init", "sim"])
@pytest.mark.parametrize("fixture", FIXTURES)
def test_gpu_full_run(fixture, kind):
    from betse_b200 import network as netlib
    from betse_b200.engine import TissueEngine
    cap = util.load_golden(fixture)
    eng = TissueEngine(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0."))
    specs = util.channels_of(cap, kind)
    desc = util.networks_of(cap, kind)[0]
    comp = netlib.compile_network(desc, eng.Co, eng.M)
    eng.set_network(comp, handler=0)
    for c in specs:
        c["handler"] = 0
        c["mod_prog"] = comp["mod_index"][comp["chan_names"].index(c["name"])]
    eng.set_channels(specs, phase_init=(kind == "init"))

    def step(e):
        st = e.step(1)
        assert not (st & (3 | 16)), st

    got = _run_traces(eng, cap, kind, step, lambda e, f: e.download([f])[f])
    _check(cap, kind, got, "gpu")
    K = max(util.snap_steps(cap, kind))
    ref = util.group(cap, "%s.k%d." % (kind, K))
    out = eng.download(["cc_cells", "cc_env", "gjopen"])
    for f, a in out.items():
        assert util.rel_err(a, ref[f]) <= CONC_REL, (kind, f, util.rel_err(a, ref[f]))
    c = eng.network_state(0)
    for k, nme in enumerate(desc["species"]):
        assert util.rel_err(c[k], ref["net0.c_cells"][k]) <= CONC_REL, nme
    eng.close()
