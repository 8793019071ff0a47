"""betse_b200/channels.py restates the reference's voltage-gated channel classes as a table of
gating terms; every tabulated model is held to the reference's OWN class
(betse/science/channels/vg_*.py `_init_state` / `_calculate_state`) over a voltage sweep.
Build container only (needs /root/reference)."""
import numpy as np
import pytest

from betse_b200 import channels as ch

pytestmark = pytest.mark.reference


def _ref_class(model):
    from oracle import refshim
    refshim.bypass_science_init()
    import importlib
    M = ch.MODELS[model]
    mod = importlib.import_module("betse.science.channels." + M.get("module", ch.CLASS_OF_ION[M["ion"]]))
    return getattr(mod, model)


@pytest.mark.parametrize("model", sorted(ch.MODELS))
def test_table_matches_reference_class(model):
    vm = np.linspace(-0.120, 0.060, 721) + 1.234e-5      # avoids the removable 0/0 points of the rates
    obj = _ref_class(model)()
    import types
    obj.init(vm.copy(), types.SimpleNamespace(mem_i=np.arange(len(vm))), None, targets=None)   # cation.py:47 reads cells.mem_i
    M = ch.MODELS[model]
    if "ml" in M:
        # Morris-Lecar (vg_morrislecar.py): one gate, V = 1000*vm, P = m, its own update (update_ml)
        m0 = np.array(obj.m, dtype=float) * np.ones_like(vm)
        obj._calculate_state(vm * 1000)
        got = ch.gates(model, vm)
        for name, a, b in zip(("mInf", "mTau"), got, (obj._mInf, obj._mTau)):
            b = np.asarray(b, dtype=float) * np.ones_like(vm)
            assert np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)) < 5e-13, (model, name)
        assert np.max(np.abs(ch.initial_state(model, vm)[0] - m0)) < 1e-13
        assert obj.Phi == M["ml"]["phi"] and bool(obj.kinetic_gate) == M["ml"]["kinetic"]
        ions, perms = ch.ions_of(model)
        assert obj.time_unit == M["time_unit"] and list(obj.ions) == ions and [float(x) for x in obj.rel_perm] == perms
        assert M["mpow"] == 1 and M["hpow"] == 0
        # one time step of the class itself (run -> update_ml / m = mInf) against the table's update, and against the
        # Hodgkin-Huxley form the device runs it through (device_quantities)
        vm2 = vm[::-1].copy()
        obj.run(vm2, types.SimpleNamespace(dt=1.0e-4))
        m1, _ = ch.advance_gates(model, m0, np.ones_like(vm), vm2, 1.0e-4)
        assert np.max(np.abs(m1 - np.asarray(obj.m, dtype=float))) < 1e-15
        assert np.max(np.abs(np.asarray(obj.P, dtype=float) - m1)) < 1e-15
        U = vm2 * 1000
        q = ch.device_quantities(model)
        mInf, tau = ch.quantity(q[0], U), ch.quantity(q[1], U)
        dtu = 1.0e-4 * M["time_unit"]
        assert np.max(np.abs((tau * m0 + dtu * mInf) / (tau + dtu) - m1)) < 1e-14
        return
    V = vm * 1000 + obj.v_corr
    m0, h0 = np.array(obj.m, dtype=float) * np.ones_like(vm), np.array(obj.h, dtype=float) * np.ones_like(vm)
    obj._calculate_state(V)
    ref = [np.asarray(x, dtype=float) * np.ones_like(vm) for x in (obj._mInf, obj._mTau, obj._hInf, obj._hTau)]
    got = ch.gates(model, vm)
    for name, a, b in zip(("mInf", "mTau", "hInf", "hTau"), got, ref):
        err = np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
        assert err < 5e-13, (model, name, err)
    gm, gh = ch.initial_state(model, vm)
    assert np.max(np.abs(gm - m0)) < 1e-13 and np.max(np.abs(gh - h0)) < 1e-13
    M = ch.MODELS[model]
    assert float(obj._mpower) == M["mpow"] and float(obj._hpower) == M["hpow"]
    ions, perms = ch.ions_of(model)
    assert obj.time_unit == M["time_unit"] and list(obj.ions) == ions and [float(x) for x in obj.rel_perm] == perms
