"""Ensembles of small tissues (last bullet): B independent members advanced by one CUDA graph
(betse_ensemble_step) must each equal the same tissue stepped alone, bit for bit — and members with different
parameters must differ from each other."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

FIELDS = ["cc_cells", "cc_env", "vm", "gjopen"]


def _members(name, B):
    cap = util.load_golden(name)
    mesh, p, s0 = util.group(cap, "cells."), util.group(cap, "sim.p."), util.group(cap, "sim.s0.")
    out = []
    for j in range(B):
        s = dict(s0)
        s["Dm_cells"] = np.asarray(s0["Dm_cells"]) * (1.0 + 0.05 * j)          # a parameter ensemble: membrane permeabilities
        out.append((mesh, p, s))
    return out


@pytest.mark.parametrize("name", ["mammal_ecm", "mammal_noecm"])
@pytest.mark.parametrize("steps", [1, 7, 23])
def test_members_equal_solo_runs_bit_exactly(name, steps):
    from betse_b200.engine import EnsembleEngine, TissueEngine
    B = 4
    members = _members(name, B)
    ens = EnsembleEngine(members=members)
    ens.update_V()
    st = ens.step(steps)
    assert not (st & 3)
    got = ens.download(FIELDS)
    ens.close()
    for j, (mesh, p, s) in enumerate(members):
        solo = TissueEngine(mesh, p, s)
        solo.update_V()
        assert not (solo.step(steps) & 3)
        ref = solo.download(FIELDS)
        solo.close()
        for f in FIELDS:
            assert np.array_equal(got[f][j], ref[f]), (f, j)
    assert not np.array_equal(got["vm"][0], got["vm"][1])


def test_ensemble_of_synthetic_tissues_profile_and_lockstep():
    from betse_b200 import synth
    from betse_b200.engine import EnsembleEngine, TissueEngine
    mesh, p, state = synth.make_tissue(10_000)
    ens = EnsembleEngine(mesh, p, state, replicas=3)
    ens.update_V()
    ens.step(3)
    ms, _ = ens.profile(20)
    assert ms > 0
    got = ens.download(["vm", "cc_cells"])
    ens.close()
    solo = TissueEngine(mesh, p, state)
    solo.update_V()
    solo.step(3 + 10 * 2 + 20)          # profile(20): two capture launches of 10 + the 20 timed steps
    ref = solo.download(["vm", "cc_cells"])
    solo.close()
    for j in range(3):
        assert np.array_equal(got["vm"][j], ref["vm"])
        assert np.array_equal(got["cc_cells"][j], ref["cc_cells"])
