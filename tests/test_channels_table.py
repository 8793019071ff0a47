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
