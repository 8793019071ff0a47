"""Host-side logic of the domain decomposition (betse_b200/partition.py), on CPU.

The exchange plans are checked with a NumPy mirror of csrc/xchg.cu's index arithmetic on arrays
labelled with GLOBAL ids, in one process and across two `gloo` ranks (the N>1 path without a GPU).
"""
import os
import socket

import numpy as np
import pytest

from betse_b200 import partition as pt
from betse_b200 import synth


def _tissue(n=1500):
    return synth.make_tissue(n)


def _labelled(parts, I=3):
    """Per-rank arrays whose OWNED entries carry global ids (ghost / halo entries = -1)."""
    out = []
    for p in parts:
        Cl = len(p.cells_local)
        El = (p.row_hi - p.row_lo) * p.nx
        cc_mid = -np.ones((I, Cl))
        vm = -np.ones(Cl)
        cc_mid[:, :p.Co] = p.own_cells[None, :] + 1e7 * np.arange(I)[:, None]
        vm[:p.Co] = p.own_cells
        flux = -np.ones((p.part["n_flux_slots"], I))
        flux[:p.Mo] = p.own_mems[:, None] + 1e7 * np.arange(I)[None, :]
        cc_env = -np.ones((I, El))
        v_raw = -np.ones(El)
        own = slice((p.a - p.row_lo) * p.nx, (p.b - p.row_lo) * p.nx)
        gsq = np.arange(p.a * p.nx, p.b * p.nx)
        cc_env[:, own] = gsq[None, :] + 1e7 * np.arange(I)[:, None]
        v_raw[own] = gsq
        out.append({"cc_mid": cc_mid, "vm": vm, "flux": flux, "cc_env": cc_env, "v_raw": v_raw})
    return out


def payload(p, side, arr):
    """What rank p pushes to its neighbour on `side` (mirror of k_xchg's PUSH loops)."""
    pl = p.plans[side]
    n, s0, _ = pl["cc_rows"]
    nv, v0, _ = pl["v_rows"]
    nx = p.nx
    return {"cells": arr["cc_mid"][:, pl["send_cells"]], "vm": arr["vm"][pl["send_cells"]],
            "flux": arr["flux"][pl["send_flux"]],
            "cc": arr["cc_env"][:, s0 * nx:(s0 + n) * nx], "v": arr["v_raw"][v0 * nx:(v0 + nv) * nx]}


def apply(pl, nx, pay, arr):
    """Write a neighbour's payload into this rank's arrays at the plan's receive offsets."""
    g0, s0 = pl["recv_cell0"], pl["recv_slot0"]
    n = pay["vm"].shape[0]
    arr["cc_mid"][:, g0:g0 + n] = pay["cells"]
    arr["vm"][g0:g0 + n] = pay["vm"]
    arr["flux"][s0:s0 + pay["flux"].shape[0]] = pay["flux"]
    ncc, _, d0 = pl["cc_rows"]
    nv, _, dv0 = pl["v_rows"]
    arr["cc_env"][:, d0 * nx:(d0 + ncc) * nx] = pay["cc"]
    arr["v_raw"][dv0 * nx:(dv0 + nv) * nx] = pay["v"]


def check_rank(p, arr, mesh, I=3):
    """After the exchange every ghost cell, remote flux slot and halo row holds its global id."""
    nx = p.nx
    assert np.array_equal(arr["vm"], p.cells_local.astype(float))
    assert np.array_equal(arr["cc_mid"][1], p.cells_local + 1e7)
    # env square -> slot CSR: exactly the global membranes of each owned square, ascending
    m2e = np.asarray(mesh["map_mem2ecm"])
    ptr, idx = p.part["ecm_slot_ptr"], p.part["ecm_slot_idx"]
    order = np.argsort(m2e, kind="stable")
    gptr = np.concatenate(([0], np.cumsum(np.bincount(m2e, minlength=p.nx * p.ny))))
    for gy in (p.a, (p.a + p.b) // 2, p.b - 1):
        for x in range(0, nx, max(1, nx // 17)):
            k = (gy - p.row_lo) * nx + x
            got = arr["flux"][idx[ptr[k]:ptr[k + 1]], 0]
            want = order[gptr[gy * nx + x]:gptr[gy * nx + x + 1]]
            assert np.array_equal(got, want.astype(float)), (p.rank, gy, x)
    # halo rows: cc_env over G+2 rows, v_raw over G+6 rows on each interior side
    for lo_g, hi_g, name in ((p.a - p.G - pt.CC_HALO, p.b + p.G + pt.CC_HALO, "cc"),
                             (p.a - p.G - pt.V_HALO, p.b + p.G + pt.V_HALO, "v")):
        lo_g, hi_g = max(lo_g, 0), min(hi_g, p.ny)
        loc = slice((lo_g - p.row_lo) * nx, (hi_g - p.row_lo) * nx)
        want = np.arange(lo_g * nx, hi_g * nx).astype(float)
        got = arr["cc_env"][0, loc] if name == "cc" else arr["v_raw"][loc]
        assert np.array_equal(got, want), (p.rank, name)


@pytest.mark.parametrize("R", [1, 2, 3, 5])
def test_partition_is_a_partition(R):
    mesh, p, st = _tissue()
    parts = pt.partition(mesh, p, st, R)
    C, M = len(mesh["cell_vol"]), len(mesh["mem_sa"])
    assert np.array_equal(np.sort(np.concatenate([q.own_cells for q in parts])), np.arange(C))
    assert np.array_equal(np.sort(np.concatenate([q.own_mems for q in parts])), np.arange(M))
    nn = np.asarray(mesh["nn_i"])
    m2c = np.asarray(mesh["mem_to_cells"])
    for q in parts:
        # local partner decoding gives back the global partner cell
        nl = q.mesh["nn_i"]
        own = nl >= 0
        assert np.array_equal(q.own_mems[nl[own]], nn[q.own_mems][own])
        gh = -(nl[~own] + 2)
        assert np.all(gh >= q.Co)
        assert np.array_equal(q.cells_local[gh], m2c[nn[q.own_mems][~own]])
        assert q.mesh["cell_mem_ptr"][-1] == q.Mo
        yi, ya, yf = q.rows["yi"], q.rows["ya"], q.rows["yf"]
        assert 0 <= yf[0] <= yi[0] <= ya[0] < ya[1] <= yi[1] <= yf[1] <= q.row_hi - q.row_lo
    # gathered state equals the input
    got = pt.gather(parts, [{"cc_cells": q.state["cc_cells"], "gjopen": q.state["gjopen"],
                             "cc_env": q.state["cc_env"]} for q in parts])
    assert np.array_equal(got["cc_cells"], st["cc_cells"])
    assert np.array_equal(got["gjopen"], st["gjopen"])
    assert np.array_equal(got["cc_env"], np.asarray(st["cc_env"]).reshape(got["cc_env"].shape))


@pytest.mark.parametrize("R", [2, 4])
def test_exchange_plans_deliver_global_ids(R):
    mesh, p, st = _tissue(2500)
    parts = pt.partition(mesh, p, st, R)
    arrs = _labelled(parts)
    for q in parts:
        for side, pl in q.plans.items():
            nbr = parts[pl["rank"]]
            apply(pl, q.nx, payload(q, side, arrs[q.rank]), arrs[nbr.rank])
    for q in parts:
        check_rank(q, arrs[q.rank], mesh)


def test_too_thin_strips_are_refused():
    mesh, p, st = _tissue(400)
    with pytest.raises(pt.BetseB200Error):
        pt.partition(mesh, p, st, 8)


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        mesh, p, st = _tissue(2500)
        parts = pt.partition(mesh, p, st, world)      # every rank derives the same plan
        me = parts[rank]
        arr = _labelled(parts)[rank]
        # push to each neighbour, receive from each neighbour (plans are symmetric: the neighbour's
        # plan towards me tells where its payload lands)
        for side, pl in sorted(me.plans.items()):
            pay = payload(me, side, arr)
            keys = sorted(pay)
            reqs = [dist.isend(torch.from_numpy(np.ascontiguousarray(pay[k])), pl["rank"]) for k in keys]
            nbr = parts[pl["rank"]]
            npl = nbr.plans[1 - side]
            shapes = {k: v.shape for k, v in payload(nbr, 1 - side, _labelled(parts)[nbr.rank]).items()}
            got = {}
            for k in keys:
                t = torch.empty(shapes[k], dtype=torch.float64)
                dist.recv(t, pl["rank"])
                got[k] = t.numpy()
            for r_ in reqs:
                r_.wait()
            apply(npl, me.nx, got, arr)
        check_rank(me, arr, mesh)
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, "FAIL %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in procs]
    for pr in procs:
        pr.join(60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_ownership_matches_partition():
    """partition.ownership (what the N-GPU drop-in uses to assemble global arrays) agrees with the full partition."""
    from betse_b200 import synth
    from betse_b200.partition import ownership, partition
    mesh, p, st = synth.make_tissue(6000)
    parts = partition(mesh, p, st, 3)
    own = ownership(mesh, 3)
    for a, b in zip(parts, own):
        assert np.array_equal(a.own_cells, b.own_cells) and np.array_equal(a.own_mems, b.own_mems)
        assert (a.a, a.b, a.row_lo, a.row_hi, a.Co, a.Mo) == (b.a, b.b, b.row_lo, b.row_hi, b.Co, b.Mo)


def test_live_scheduled_follows_the_event_options():
    """simloop._live_scheduled: only what the phase's events can move (tishandler.py:744-876) is compared per step."""
    import types
    from betse_b200 import simloop
    off = {k: 0 for k in ("K_env", "Cl_env", "Na_env", "T_change", "gj_block", "NaKATP_block")}
    soff = {k: 0 for k in ("Na_mem", "K_mem", "Cl_mem", "Ca_mem", "pressure", "ecmJ", "cuts")}
    p = types.SimpleNamespace(global_options=dict(off), scheduled_options=dict(soff))
    assert simloop._live_scheduled(p) == []
    p.global_options["gj_block"] = [1.0, 2.0, 0.5]
    p.scheduled_options["K_mem"] = [1.0, 2.0, 0.5, 10.0, ["Spot"]]
    assert simloop._live_scheduled(p) == ["Dm_cells", "gj_block"]
    p.global_options.update(K_env=[1, 2, 3, 4], T_change=[1, 2, 3, 4], NaKATP_block=[1, 2, 3])
    p.scheduled_options["ecmJ"] = [1, 2, 3, ["Spot"], 1.0]
    assert simloop._live_scheduled(p) == ["Dm_cells", "D_env", "gj_block", "NaKATP_block", "c_env_bound", "T"]
    assert simloop._live_scheduled(types.SimpleNamespace()) == list(simloop._SCHEDULED)      # unknown parameter object: all
