"""Drivers of a domain-decomposed tissue: strips of env-grid rows, one ``betse_ctx``
per strip, halo exchange by peer stores into the neighbours' windows (csrc/xchg.cu).

* :class:`DistributedStrips` — production: one process per GPU (``torchrun``); the window handles
  travel once through ``torch.distributed`` (plumbing), after that a timestep is ONE CUDA-graph
  launch per rank and no host or NCCL call is on the data path.
* :class:`LocalStrips` — every strip in this process on one device, the exchange points driven
  from the host in lock-step; used by the parity tests (a 1-GPU box can check any strip count
  bit-exactly against the undivided run).
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import BetseB200Error
from .engine import TissueEngine
from .partition import gather, partition


def _rank_engine(part, params, device):
    eng = TissueEngine(part.mesh, params, part.state, device=device, partition=part.part)
    eng.set_row_ranges(part.rows["yi"], part.rows["ya"], part.rows["yf"])
    return eng


def _info_to_bytes(w):
    return bytes(memoryview(w))


def _info_from_bytes(b):
    w = capi.WindowInfo()
    C.memmove(C.addressof(w), b, C.sizeof(w))
    return w


class LocalStrips:
    def __init__(self, mesh, params, state, R, device=0, bounds=None):
        self.parts = partition(mesh, params, state, R, bounds=bounds)
        self.engines = [_rank_engine(p, params, device) for p in self.parts]
        infos = [e.window() for e in self.engines]
        for p, e in zip(self.parts, self.engines):
            for side, plan in p.plans.items():
                e.attach_neighbor(side, infos[plan["rank"]], plan, same_process=True)
        self.R = R

    def _sync(self):
        st = 0
        for e in self.engines:
            st |= e.sync()
        return st

    def _exchange(self, which, buf_next):
        if self.R == 1:
            return
        self._sync()
        for e in self.engines:
            e.exchange(which, buf_next, capi.XCHG_PUSH)
        self._sync()
        for e in self.engines:
            e.exchange(which, buf_next, capi.XCHG_WAIT)

    def update_V(self):
        for e in self.engines:
            e.update_V_phase(0)
        self._exchange(capi.XCHG_X1, False)
        self._exchange(capi.XCHG_X2, False)
        for e in self.engines:
            e.update_V_phase(1)
        return self._sync()

    def step(self, n=1, diag=False):
        st = 0
        for k in range(n):
            d = diag and k == n - 1
            for e in self.engines:
                e.step_phase(0, d)
            self._exchange(capi.XCHG_X1, True)
            for e in self.engines:
                e.step_phase(1, d)
            self._exchange(capi.XCHG_X2, True)
            for e in self.engines:
                e.step_phase(2, d)
            st |= self._sync()
        return st

    def download(self, fields):
        return gather(self.parts, [e.download(fields) for e in self.engines])

    def close(self):
        for e in self.engines:
            e.close()


class DistributedStrips:
    """This process is rank ``dist.get_rank()`` of ``dist.get_world_size()`` strips.  Every rank
    passes the same global mesh / state (synthetic tissues are rebuilt from the seed on each rank;
    a loaded tissue is read by each rank) and keeps only its strip."""

    def __init__(self, mesh, params, state, device, dist, bounds=None):
        self.dist = dist
        self.rank, self.R = dist.get_rank(), dist.get_world_size()
        self.part = partition(mesh, params, state, self.R, bounds=bounds, only=self.rank)[self.rank]
        self.engine = _rank_engine(self.part, params, device)
        mine = _info_to_bytes(self.engine.window())
        allinfo = [None] * self.R
        dist.all_gather_object(allinfo, mine)
        for side, plan in self.part.plans.items():
            self.engine.attach_neighbor(side, _info_from_bytes(allinfo[plan["rank"]]), plan, same_process=False)
        dist.barrier()

    def update_V(self):
        self.engine.update_V()

    def step(self, n=1, diag=False):
        st = self.engine.step(n, diag=diag)
        if st & capi.STATUS_XCHG_TIMEOUT:
            raise BetseB200Error("halo exchange timed out waiting for a neighbouring rank")
        return st

    def set_field(self, name, value):
        """A scheduled quantity of the GLOBAL tissue (what TissueHandler.fire_events rewrote): scalars and per-ion vectors
        go to the rank as they are, per-membrane / per-env-square arrays are cut to this rank's strip first."""
        v = np.asarray(value)
        P = self.part
        if v.ndim >= 1 and v.shape[-1] == P.g2l_m.shape[0]:                     # [.., M]
            v = v[..., P.own_mems]
        elif v.ndim >= 1 and v.shape[-1] == P.ny * P.nx:                         # [.., E]
            v = v[..., P.row_lo * P.nx:P.row_hi * P.nx]
        self.engine.set_field(name, v)

    def profile(self, n):
        return self.engine.profile(n)

    def download_local(self, fields, pinned=False):
        """This rank's strip of the named fields.  ``pinned``: views of the engine's page-locked staging, overwritten by
        the next pinned download (for consumers that copy what they keep, like write2storage)."""
        return self.engine.download(fields, pinned=pinned)

    def close(self):
        self.dist.barrier()
        self.engine.close()


def verify_strips(dist, device, cells=60_000, steps=12, fields=("cc_cells", "cc_env", "vm", "gjopen", "E_env_x", "v_env")):
    """Multi-process parity check, run by every rank of an initialised process group: one synthetic tissue stepped as
    ``world`` strips (halo exchange over CUDA-IPC peer stores) must equal the same tissue stepped undivided on rank 0's
    GPU, BIT FOR BIT.  Returns the verdict dict on rank 0 (None elsewhere).  bench.py runs it before timing a decomposed
    tissue; tools/check_multigpu.py is the stand-alone form."""
    from . import synth
    from .partition import gather, partition
    mesh, p, st = synth.make_tissue(cells)
    ds = DistributedStrips(mesh, p, st, device, dist)
    ds.update_V()
    status = ds.step(steps)
    mine = ds.download_local(list(fields))
    allf = [None] * dist.get_world_size()
    dist.all_gather_object(allf, mine)
    ds.close()
    out = None
    if dist.get_rank() == 0:
        parts = partition(mesh, p, st, dist.get_world_size())
        got = gather(parts, allf)
        eng = TissueEngine(mesh, p, st, device=device)
        eng.update_V()
        s0 = eng.step(steps)
        ref = eng.download(list(fields))
        eng.close()
        worst, ok = {}, True
        for f in fields:
            a, b = got[f].reshape(ref[f].shape), ref[f]
            same = bool(np.array_equal(a, b))
            worst[f] = 0.0 if same else float(np.max(np.abs(a - b)))
            ok &= same
        out = {"check": "strips == single domain (bit-exact)", "ok": ok, "world": dist.get_world_size(),
               "cells": len(mesh["cell_vol"]), "steps": steps, "status": [int(status), int(s0)], "max_abs_diff": worst}
    dist.barrier()
    return out
