"""Domain decomposition on the GPU: a tissue cut into R strips (one betse_ctx per strip, halo
exchange by stores into the neighbours' windows, csrc/xchg.cu) must reproduce the undivided run
BIT-EXACTLY — every sum is taken in the same order (partition.py orders each env square's flux
slots by global membrane index).  All strips live on one device here, so a 1-GPU box checks any
strip count; the multi-process / NVLink transport is exercised by `bench.py --gpus N`."""
import numpy as np
import pytest

from tests import util

pytestmark = pytest.mark.gpu

FIELDS = ["cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "rho_cells", "E_env_x", "E_env_y", "v_env", "rho_env"]


def _single(mesh, p, st, steps):
    from betse_b200.engine import TissueEngine
    eng = TissueEngine(mesh, p, st)
    eng.update_V()
    status = eng.step(steps)
    out = eng.download(FIELDS)
    eng.close()
    return status, out


@pytest.mark.parametrize("n_cells,R,steps", [(20_000, 2, 12), (20_000, 3, 7), (60_000, 4, 5)])
def test_strips_equal_single_domain(n_cells, R, steps):
    from betse_b200 import synth
    from betse_b200.strips import LocalStrips
    mesh, p, st = synth.make_tissue(n_cells)
    s0, ref = _single(mesh, p, st, steps)
    ls = LocalStrips(mesh, p, st, R)
    assert ls.update_V() == 0
    s1 = ls.step(steps)
    got = ls.download(FIELDS)
    ls.close()
    assert not (s0 & 11) and not (s1 & 11)
    for f in FIELDS:
        assert np.array_equal(got[f].reshape(ref[f].shape), ref[f]), (f, float(np.max(np.abs(got[f].reshape(ref[f].shape) - ref[f]))))


def test_strips_on_reference_mesh_match_golden():
    """Ragged reference-built mesh (3-7 membranes per cell, 24x25 grid) in 2 strips against the
    real reference's recorded outputs."""
    from betse_b200.strips import LocalStrips
    cap = util.load_golden("mammal_ecm")
    kind = "sim"
    mesh, p, s0 = util.group(cap, "cells."), util.group(cap, kind + ".p."), util.group(cap, kind + ".s0.")
    ls = LocalStrips(mesh, p, s0, 2)
    K = 5
    for n in range(K):
        assert not util.group(cap, "%s.sched.k%d." % (kind, n + 1)), "fixture has scheduled changes"
        st = ls.step(1)
        assert not (st & 11)
    ref = util.group(cap, "%s.k%d." % (kind, K))
    got = ls.download(["cc_cells", "cc_env", "vm", "gjopen"])
    ls.close()
    tol = util.gpu_tolerances(cap, kind, ref)
    for f, a in got.items():
        err = float(np.max(np.abs(a.reshape(np.shape(ref[f])) - ref[f])))
        assert err <= tol[f], (f, err, tol[f])


def test_multiprocess_strips_over_ipc():
    """One process per GPU, CUDA-IPC windows over NVLink (tools/check_multigpu.py); needs >= 2 GPUs."""
    import os
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = min(n, 4)
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", "29533",
                          os.path.join(root, "tools", "check_multigpu.py"), "--cells", "60000", "--steps", "12"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    assert '"ok": true' in res.stdout


def test_dropin_loop_over_two_gpus_matches_reference():
    """run_sim_core_loop inside a torchrun process group: the unmodified reference's seed/init/sim with the tissue stepped
    as strips, one per GPU, against the reference's own loop (tools/check_dropin_multigpu.py); needs >= 2 GPUs and the
    reference tree (baseline/_ref)."""
    import os
    import subprocess
    import sys
    import torch
    from oracle import refshim
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    if not refshim.reference_available():
        pytest.skip("reference tree absent: neither /root/reference nor baseline/_ref (run tools/install_reference.py)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29544",
                          os.path.join(root, "tools", "check_dropin_multigpu.py")],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-3000:]
    assert '"ok": true' in res.stdout
