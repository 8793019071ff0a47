"""Pins oracle/betse_oracle.py to the REAL reference: tests/golden/*.npz were produced by
running the unmodified reference (tests/golden/make_golden.py); the oracle must reproduce
every recorded field after every recorded number of timesteps."""
import numpy as np
import pytest

from oracle.betse_oracle import OracleSim
from tests import util


@pytest.mark.parametrize("name", [n for n in util.GOLDEN if not n.startswith("default_try")])   # full run: tests/test_full_run.py
@pytest.mark.parametrize("kind", ["init", "sim"])
def test_oracle_reproduces_reference(name, kind):
    cap = util.load_golden(name)
    o = OracleSim(util.mesh_of(cap, kind), util.group(cap, kind + ".p."),
                  util.group(cap, kind + ".s0."), channels=util.channels_of(cap, kind), phase_init=(kind == "init"),
                  networks=util.networks_of(cap, kind), net_handlers=util.network_handlers(cap, kind))
    n = 0
    checked = 0
    for K in util.snap_steps(cap, kind):
        while n < K:
            util.apply_schedule(o, cap, kind, n + 1)
            util.set_time(o, cap, kind, n + 1)
            o.step()
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        polar = float(cap[kind + ".p.cell_polarizability"]) != 0.0
        for f in util.STATE + util.ENV_STATE + util.DIAG:
            if f not in ref or not hasattr(o, f):
                continue
            if not int(cap[kind + ".p.is_ecm"]) and f in util.ENV_STATE:
                continue
            tol = 1e-12
            if polar and f in ("E_cell_x", "E_cell_y", "Emc"):
                tol = 1e-6  # Σ(vm - vm_ave)·n̂·sa: catastrophic cancellation (see util.scale_of)
            if f == "B_field":
                tol = 1e-9
            err = util.rel_err(getattr(o, f), ref[f], util.scale_of(f, ref))
            assert err < tol, (name, kind, K, f, err)
            checked += 1
        for k, c in enumerate(o.channels):          # gate states of the channel library
            for f in ("m", "h", "P"):
                key = "chan%d.%s" % (k, f)
                if key in ref and not (kind == "init" and not c["init_active"]):
                    assert util.rel_err(c[f], ref[key]) < 1e-12, (name, kind, K, key)
                    checked += 1
        if name == "mammal_ecm_net_tj" and "TJ_modulator" in ref:                 # what the tight-junction modulators left (networks.py:3301-3317)
            assert util.rel_err(o.TJ_modulator, ref["TJ_modulator"]) < 1e-12, (name, kind, K)
            assert float(np.ptp(ref["TJ_modulator"])) > 0.5
            checked += 1
        for h, net in zip(util.network_handlers(cap, kind), o.networks):        # network substances, rates, channel DChan (networks.py:2805-2982, 3164)
            want = ref["net%d.c_cells" % h]
            for k, nme in enumerate(net.species):
                assert util.rel_err(net.c[nme], want[k]) < 1e-12, (name, kind, K, nme)
                checked += 1
            if "net%d.c_mems" % h in ref:            # membrane values under intracellular transport (networks.py:5727-5795)
                for k, nme in enumerate(net.species):
                    assert util.rel_err(net.cmem[nme], ref["net%d.c_mems" % h][k]) < 1e-12, (name, kind, K, nme, "mems")
                    checked += 1
            if "net%d.c_env" % h in ref:             # membrane / extracellular legs of molecule_mover (sim_toolbox.py:909-1153)
                for k, nme in enumerate(net.species):
                    if net.env_on[k]:
                        assert util.rel_err(net.c_env[nme], ref["net%d.c_env" % h][k]) < 1e-12, (name, kind, K, nme, "env")
                        checked += 1
            if "net%d.reaction_rates" % h in ref:
                nrx = ref["net%d.reaction_rates" % h].shape[0]
                assert util.rel_err(net.rates[-nrx:], ref["net%d.reaction_rates" % h]) < 1e-12
            for k, c in enumerate(o.channels):
                key = "chan%d.DChan" % k
                if key in ref and "DChan" in c and not (kind == "init" and not c["init_active"]):
                    assert util.rel_err(c["DChan"], ref[key]) < 1e-12, (name, kind, K, key)
                    checked += 1
    assert checked > 20


@pytest.mark.parametrize("kind", ["init", "sim"])
@pytest.mark.parametrize("fixture", ["fast_basic", "fast_mammal_noecm", "fast_chan"])
def test_fast_solver_oracle_matches_reference(fixture, kind):
    """The equivalent-circuit solver (sim.py:1454-1640) restated in oracle.OracleFastSim against the real reference;
    fast_chan: with the voltage-gated channels of run_fast_loop_channels (networks.py:3217-3280)."""
    from oracle.betse_oracle import OracleFastSim
    cap = util.load_golden(fixture)
    s0 = util.group(cap, kind + ".s0.")
    o = OracleFastSim(util.mesh_of(cap, kind), util.group(cap, kind + ".p."), s0, channels=util.channels_of(cap, kind),
                      phase_init=(kind == "init"), fast_consts=util.fast_consts_of(cap, kind))
    n = 0
    for K in util.snap_steps(cap, kind):
        while n < K:
            o.step()
            n += 1
        ref = util.group(cap, "%s.k%d." % (kind, K))
        for f in ("vm_ave", "vm", "gjopen", "vgj", "Jn", "J_cell_x", "J_cell_y", "E_cell_x", "E_cell_y", "Emx", "Emy") + \
                (("extra_J_mem",) if o.channels else ()):
            if f in ref:
                a, r = getattr(o, f), ref[f]
                assert np.max(np.abs(a - r)) <= 1e-12 * max(np.max(np.abs(r)), 1e-300), (kind, K, f)
        for k, c in enumerate(o.channels):
            if kind == "init" and not c["init_active"]:
                continue
            for f in ("m", "h", "P", "flux"):
                r = ref.get("chan%d.%s" % (k, f))
                if r is not None and f in c:
                    assert np.max(np.abs(c[f] - r)) <= 1e-12 * max(np.max(np.abs(r)), 1e-300), (kind, K, k, f)
