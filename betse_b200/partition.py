"""Spatial domain decomposition of a tissue into strips of env-grid rows.

Host-side integer work only (NumPy): given the global mesh / state dicts (the ``Cells`` and
``Simulator`` attributes the loop consumes) it produces, per rank,

* a local mesh: owned cells in global order followed by ghost cells (gap-junction partners owned
  by the strip below, then above), the membranes of the owned cells, a window of env-grid rows
  ``[lo, hi)`` around the owned rows ``[a, b)`` and the env-square -> flux-slot CSR of the owned
  squares, whose entries are ordered by GLOBAL membrane index so that every sum is taken in the
  same order as on one GPU (results are bit-identical for any number of ranks);
* the local state (ghost cells and halo rows hold the neighbours' initial values);
* the exchange plan towards each neighbour (include/betse_b200.h ``betse_neighbor``).

What couples strips is exactly what the reference's index arrays couple: ``cells.nn_i``
(cells.py:1498-1533), ``cells.map_mem2ecm`` (cells.py:1758) and the grid stencils of
``update_ecm`` (sim.py:2209-2254, radius 2) and of ``get_current``'s env field
(ion_current.py:101-109: 9-tap Gaussian, radius 4, + gradient, radius 1).
"""
import numpy as np

from .capi import BetseB200Error

CC_HALO = 2      # update_ecm: flux at +-1 needs the gradient at +-1, i.e. concentrations at +-2
V_HALO = 6       # E rows [a-G-1, b+G+1) need v_raw at +-5 around them


def _bflags_bool(mesh, M):
    b = np.asarray(mesh["bflags_mems"])
    if b.dtype == np.bool_ and b.size == M:
        return b.copy()
    out = np.zeros(M, dtype=bool)
    out[b.astype(np.int64)] = True
    return out


def strip_bounds(mesh, R):
    """Row boundaries [a_0=0, a_1, ..., a_R=ny] balancing membranes (+ a grid term) per strip."""
    ny, nx = (int(x) for x in mesh["grid_shape"])
    mrow = np.asarray(mesh["map_mem2ecm"], dtype=np.int64) // nx
    w = np.bincount(mrow, minlength=ny).astype(float) + 0.3 * nx
    cw = np.concatenate(([0.0], np.cumsum(w)))
    b = [0]
    for r in range(1, R):
        b.append(int(np.searchsorted(cw, cw[-1] * r / R)))
    b.append(ny)
    return np.asarray(b, dtype=np.int64)


class RankPart:
    """Everything rank ``r`` needs: ``mesh`` / ``state`` dicts for TissueEngine, ``part`` (the
    betse_mesh decomposition fields), ``rows`` (kernel row ranges), ``plans`` {side: plan} and the
    global index lists used to gather results."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def partition(mesh, params, state, R, bounds=None, only=None):
    """-> list of R ``RankPart``.  ``only=r``: build rank r alone (what a rank of a multi-process run needs: its own part
    and, for the exchange plans, the sizes of its two neighbours); the other entries of the list are None / size stubs."""
    if R < 1:
        raise ValueError("R must be >= 1")
    need = list(range(R)) if only is None else [r for r in (only - 1, only, only + 1) if 0 <= r < R]
    ny, nx = (int(x) for x in mesh["grid_shape"])
    ptr = np.asarray(mesh["cell_mem_ptr"], dtype=np.int64)
    m2c = np.asarray(mesh["mem_to_cells"], dtype=np.int64)
    nn = np.asarray(mesh["nn_i"], dtype=np.int64)
    m2e = np.asarray(mesh["map_mem2ecm"], dtype=np.int64)
    C, M = len(ptr) - 1, len(m2c)
    if not bool(np.asarray(params["is_ecm"]).item() if np.ndim(params["is_ecm"]) == 0 else params["is_ecm"]):
        raise BetseB200Error("domain decomposition needs extracellular spaces; no-ECM tissues run as replicas")
    if float(np.asarray(params.get("sharpness", 1.0))) < 1.0:
        raise BetseB200Error("domain decomposition does not support the env smoothing pass (sharpness < 1)")
    if "Phi_b" in state and np.any(np.asarray(state["Phi_b"]) != 0):
        raise BetseB200Error("domain decomposition does not support a boundary-voltage potential")
    if "map_cell2ecm" in mesh:
        crow = np.asarray(mesh["map_cell2ecm"]).astype(np.int64) // nx
    else:   # row of the cell's first membrane
        crow = (m2e // nx)[ptr[:-1]]
    mrow = m2e // nx
    bounds = strip_bounds(mesh, R) if bounds is None else np.asarray(bounds, dtype=np.int64)
    if len(bounds) != R + 1 or bounds[0] != 0 or bounds[-1] != ny or np.any(np.diff(bounds) <= 0):
        raise ValueError("bad strip bounds %r" % (bounds,))
    owner_c = np.searchsorted(bounds[1:], crow, side="right")
    owner_m = owner_c[m2c]                       # rank that steps the membrane
    env_m = np.searchsorted(bounds[1:], mrow, side="right")   # rank that owns its env square
    out_m = np.nonzero(env_m != owner_m)[0]     # membranes whose env square lies in a neighbour's strip: the strip edges only
    if np.any(np.abs(owner_m[out_m] - env_m[out_m]) > 1):
        raise BetseB200Error("a membrane maps to an env square two strips away: strips are too thin")
    a_m, b_m = bounds[owner_m[out_m]], bounds[owner_m[out_m] + 1]
    # reach of membranes outside their strip (zero for every membrane inside it)
    G = int(max(0, np.max(np.maximum(a_m - mrow[out_m], mrow[out_m] - (b_m - 1))))) if len(out_m) else 0
    H = G + V_HALO
    if R > 1 and np.min(np.diff(bounds)) < H:
        raise BetseB200Error("strips of %d rows are thinner than the %d-row halo; use fewer ranks"
                             % (int(np.min(np.diff(bounds))), H))
    pc = m2c[nn]                                 # partner cell of every membrane
    bfl = _bflags_bool(mesh, M)
    counts = np.diff(ptr)
    msa_env = np.asarray(mesh["memSa_per_envSquare"], dtype=float) if "memSa_per_envSquare" in mesh else None
    memsa_mean = float(msa_env[m2e].mean()) if msa_env is not None else 1.0

    # ---- pass 1: ownership lists of every rank
    own_cells, own_mems, ghosts, remote_in = {}, {}, {}, {}
    for r in need:
        oc = np.nonzero(owner_c == r)[0]
        cnt = counts[oc]
        lp = np.concatenate(([0], np.cumsum(cnt)))
        om = np.repeat(ptr[oc] - lp[:-1], cnt) + np.arange(lp[-1])
        own_cells[r] = oc
        own_mems[r] = om
        rc = pc[om]
        rem = owner_c[rc] != r
        gl = np.unique(rc[rem])
        go = owner_c[gl]
        if np.any(np.abs(go - r) > 1):
            raise BetseB200Error("a gap-junction partner lives two strips away: strips are too thin")
        ghosts[r] = (gl[go == r - 1], gl[go == r + 1])
        # membranes of the neighbours whose env square this rank owns (ascending global index)
        inc = out_m[env_m[out_m] == r]           # == nonzero((env_m == r) & (owner_m != r)), from the edge membranes
        remote_in[r] = (inc[owner_m[inc] == r - 1], inc[owner_m[inc] == r + 1])

    parts = [None] * R
    for r in need:
        a, b = int(bounds[r]), int(bounds[r + 1])
        lo, hi = max(0, a - H), min(ny, b + H)
        oc, om = own_cells[r], own_mems[r]
        Co, Mo = len(oc), len(om)
        # cells numbered along the rows (lattice-seeded meshes, the synthetic tissues) give every strip ONE range of
        # cells and membranes: plain slices then, instead of gathers over millions of indices
        om_ix = slice(int(om[0]), int(om[-1]) + 1) if Mo and int(om[-1]) - int(om[0]) + 1 == Mo else om
        oc_ix = slice(int(oc[0]), int(oc[-1]) + 1) if Co and int(oc[-1]) - int(oc[0]) + 1 == Co else oc
        g_lo, g_hi = ghosts[r]
        if only is not None and r != only:       # a neighbour: only what the exchange plan of `only` reads
            parts[r] = RankPart(rank=r, R=R, Co=Co, Mo=Mo, row_lo=lo, row_hi=hi, a=a, b=b, nx=nx, ny=ny, plans={},
                                n_ghost=(len(g_lo), len(g_hi)), n_remote=(len(remote_in[r][0]), len(remote_in[r][1])))
            continue
        cells_l = np.concatenate((oc, g_lo, g_hi))
        Cl = len(cells_l)
        g2l_c = np.full(C, -1, dtype=np.int64)
        g2l_c[cells_l] = np.arange(Cl)
        g2l_m = np.full(M, -1, dtype=np.int64)
        g2l_m[om_ix] = np.arange(Mo)
        rc = pc[om_ix]
        rem = owner_c[rc] != r
        nn_l = np.where(rem, -(g2l_c[rc] + 2), g2l_m[nn[om_ix]])
        cnt = counts[oc_ix]
        ptr_l = np.concatenate(([0], np.cumsum(cnt)))
        rows = slice(lo * nx, hi * nx)
        El = (hi - lo) * nx
        m2e_l = m2e[om_ix] - lo * nx
        if np.any(m2e_l < 0) or np.any(m2e_l >= El):
            raise BetseB200Error("internal: membrane env square outside the local window")
        # ---- env square -> flux slot CSR over the owned squares, ordered by global membrane index
        r_lo, r_hi = remote_in[r]
        mine = np.nonzero(env_m[om_ix] == r)[0]
        sq = np.concatenate((m2e_l[mine], m2e[r_lo] - lo * nx, m2e[r_hi] - lo * nx))
        gid = np.concatenate((om[mine], r_lo, r_hi))
        slot = np.concatenate((mine, Mo + np.arange(len(r_lo) + len(r_hi))))
        order = np.lexsort((gid, sq))
        slot_idx = slot[order]
        slot_ptr = np.concatenate(([0], np.cumsum(np.bincount(sq, minlength=El))))

        def cellf(name):
            return np.asarray(mesh[name])[oc_ix]

        def memf(name):
            return np.asarray(mesh[name])[om_ix]
        mesh_l = {
            "mem_to_cells": g2l_c[m2c[om_ix]], "cell_mem_ptr": ptr_l, "nn_i": nn_l, "bflags_mems": bfl[om_ix],
            "map_mem2ecm": m2e_l, "mem_sa": memf("mem_sa"), "mem_nx": memf("mem_nx"), "mem_ny": memf("mem_ny"),
            "cell_vol": cellf("cell_vol"), "cell_sa": cellf("cell_sa"), "diviterm": cellf("diviterm"),
            "num_mems": cellf("num_mems"), "delta": mesh["delta"], "gj_len": mesh["gj_len"],
            "grid_shape": np.array([hi - lo, nx]), "memsa_mean": memsa_mean,
        }
        if "ecm_vol" in mesh:
            mesh_l["ecm_vol"] = mesh["ecm_vol"]
        if "R_rads" in mesh:
            mesh_l["R_rads"] = memf("R_rads")
        if msa_env is not None:
            mesh_l["memSa_per_envSquare"] = msa_env[rows]
        if "gj_default_weights" in mesh:
            mesh_l["gj_default_weights"] = memf("gj_default_weights")
        part = {"n_cells": Cl, "n_cells_owned": Co, "n_mems_owned": Mo,
                "n_flux_slots": Mo + len(r_lo) + len(r_hi), "y0": lo, "ny_global": ny,
                "y_own0": a - lo, "y_own1": b - lo, "ecm_slot_ptr": slot_ptr, "ecm_slot_idx": slot_idx}
        krows = {"yi": (max(a - G, 0) - lo, min(b + G, ny) - lo), "ya": (a - lo, b - lo),
                 "yf": (max(a - G - 1, 0) - lo, min(b + G + 1, ny) - lo)}

        # ---- local state
        st = {}
        S = state
        I = np.asarray(S["cc_cells"]).shape[0]
        st["cc_cells"] = np.asarray(S["cc_cells"], dtype=float)[:, cells_l]
        cam = np.asarray(S["cc_at_mem"], dtype=float)
        st["cc_mid"] = cam[:, ptr[cells_l]] if cam.shape[1] == M else cam[:, cells_l]
        st["cc_env"] = np.asarray(S["cc_env"], dtype=float).reshape(I, -1)[:, rows]
        st["vm_cell"] = np.asarray(S["vm"], dtype=float)[ptr[cells_l]]
        st["gjopen"] = np.asarray(S["gjopen"], dtype=float)[om_ix]
        st["Dm_cells"] = np.asarray(S["Dm_cells"], dtype=float)[:, om_ix]
        st["D_env"] = np.asarray(S["D_env"], dtype=float).reshape(I, -1)[:, rows]
        tj = np.asarray(S.get("TJ_modulator", 1.0), dtype=float)
        st["TJ_modulator"] = tj.reshape(I, -1)[:, rows] if tj.ndim else tj
        for f in ("E_env_x", "E_env_y"):
            st[f] = np.asarray(S[f], dtype=float).reshape(-1)[rows]
        for f, idx in (("extra_rho_cells", cells_l), ("extra_J_mem", om)):
            if f in S and np.any(np.asarray(S[f]) != 0):
                st[f] = np.asarray(S[f], dtype=float)[idx]
        if "extra_rho_env" in S and np.any(np.asarray(S["extra_rho_env"]) != 0):
            st["extra_rho_env"] = np.asarray(S["extra_rho_env"], dtype=float).reshape(-1)[rows]
        for f in ("NaKATP_block", "gj_block"):
            if f in S:
                v = np.asarray(S[f], dtype=float)
                st[f] = v[om_ix] if v.ndim and v.size == M else v
        for f in ("zs", "D_free", "D_gj", "c_env_bound", "T", "ko_env", "rho_pump", "rho_channel", "bound_V"):
            if f in S:
                st[f] = S[f]
        parts[r] = (RankPart(rank=r, R=R, mesh=mesh_l, state=st, part=part, rows=krows, plans={},
                              own_cells=oc, own_mems=om, cells_local=cells_l, row_lo=lo, row_hi=hi,
                              a=a, b=b, G=G, H=H, nx=nx, ny=ny, Co=Co, Mo=Mo, g2l_c=g2l_c, g2l_m=g2l_m,
                              n_ghost=(len(g_lo), len(g_hi)), n_remote=(len(r_lo), len(r_hi))))

    # ---- pass 2: exchange plans (what rank r pushes into neighbour s)
    for r in (range(R) if only is None else [only]):
        P = parts[r]
        for side, s in ((0, r - 1), (1, r + 1)):
            if s < 0 or s >= R:
                continue
            Q = parts[s]
            # r is s's hi neighbour when s = r-1, its lo neighbour when s = r+1
            k = 1 if s == r - 1 else 0
            g_list = ghosts[s][k]                  # cells of r that are ghosts on s, ascending
            f_list = remote_in[s][k]               # membranes of r feeding s's env squares, ascending
            n_cc = min(G + CC_HALO, P.b - P.a)
            n_v = min(G + V_HALO, P.b - P.a)
            if side == 0:      # neighbour below: it needs my first owned rows [a, a+n)
                cc_g0, v_g0 = P.a, P.a
            else:              # neighbour above: my last owned rows [b-n, b)
                cc_g0, v_g0 = P.b - n_cc, P.b - n_v
            P.plans[side] = {
                "rank": s,
                "send_cells": P.g2l_c[g_list], "recv_cell0": Q.Co + (Q.n_ghost[0] if k == 1 else 0),
                "send_flux": P.g2l_m[f_list], "recv_slot0": Q.Mo + (Q.n_remote[0] if k == 1 else 0),
                "cc_rows": (n_cc, cc_g0 - P.row_lo, cc_g0 - Q.row_lo),
                "v_rows": (n_v, v_g0 - P.row_lo, v_g0 - Q.row_lo),
            }
            if np.any(P.plans[side]["send_cells"] < 0) or np.any(P.plans[side]["send_cells"] >= P.Co):
                raise BetseB200Error("internal: ghost of a neighbour is not an owned cell")
    return parts


def ownership(mesh, R, bounds=None):
    """Who owns what, for every rank, without building any rank's mesh or state (cheap: O(C + M) integer work): a list of
    light RankPart with ``own_cells``, ``own_mems`` (global indices, ascending), the owned rows ``[a, b)`` and the window
    ``[row_lo, row_hi)`` a rank's env downloads cover.  What a rank needs to assemble global arrays from every rank's strip
    (simloop's N-GPU loop); consistent with :func:`partition` by construction (same rules)."""
    ny, nx = (int(x) for x in mesh["grid_shape"])
    ptr = np.asarray(mesh["cell_mem_ptr"], dtype=np.int64)
    m2c = np.asarray(mesh["mem_to_cells"], dtype=np.int64)
    m2e = np.asarray(mesh["map_mem2ecm"], dtype=np.int64)
    crow = (np.asarray(mesh["map_cell2ecm"]).astype(np.int64) // nx) if "map_cell2ecm" in mesh else (m2e // nx)[ptr[:-1]]
    mrow = m2e // nx
    bounds = strip_bounds(mesh, R) if bounds is None else np.asarray(bounds, dtype=np.int64)
    owner_c = np.searchsorted(bounds[1:], crow, side="right")
    owner_m = owner_c[m2c]
    env_m = np.searchsorted(bounds[1:], mrow, side="right")
    out_m = np.nonzero(env_m != owner_m)[0]
    a_m, b_m = bounds[owner_m[out_m]], bounds[owner_m[out_m] + 1]
    G = int(max(0, np.max(np.maximum(a_m - mrow[out_m], mrow[out_m] - (b_m - 1))))) if len(out_m) else 0
    H = G + V_HALO
    counts = np.diff(ptr)
    out = []
    for r in range(R):
        oc = np.nonzero(owner_c == r)[0]
        cnt = counts[oc]
        lp = np.concatenate(([0], np.cumsum(cnt)))
        om = np.repeat(ptr[oc] - lp[:-1], cnt) + np.arange(lp[-1])
        a, b = int(bounds[r]), int(bounds[r + 1])
        out.append(RankPart(rank=r, R=R, own_cells=oc, own_mems=om, Co=len(oc), Mo=len(om), a=a, b=b,
                            row_lo=max(0, a - H), row_hi=min(ny, b + H), nx=nx, ny=ny, cells_local=oc))
    return out


def gather(parts, fields_per_rank):
    """Assemble global arrays from per-rank downloads ({name: array} per rank): cell fields take
    the owned cells, membrane fields the owned membranes, env fields the owned rows."""
    P0 = parts[0]
    C = sum(p.Co for p in parts)
    M = sum(p.Mo for p in parts)
    nx, ny = P0.nx, P0.ny
    out = {}
    for name in fields_per_rank[0]:
        a0 = np.asarray(fields_per_rank[0][name])
        lead = a0.shape[:-1]
        n_last = a0.shape[-1]
        from .engine import _DOWN_SHAPES          # name -> IC / IM / IE / C / M / E: sizes can coincide, names cannot
        key = {"cc_at_mem": "IM"}.get(name, _DOWN_SHAPES.get(name, ""))[-1:]
        if key == "M":
            kind, n = "M", M
        elif key == "C":
            kind, n = "C", C
        elif key == "E":
            kind, n = "E", ny * nx
        elif n_last == P0.Mo and name not in ("cc_cells", "rho_cells", "vm_ave"):
            kind, n = "M", M
        elif n_last in (P0.Co, len(P0.cells_local)):
            kind, n = "C", C
        elif n_last == (P0.row_hi - P0.row_lo) * nx:
            kind, n = "E", ny * nx
        else:
            raise ValueError("cannot classify %s with last dim %d" % (name, n_last))
        g = np.empty(lead + (n,))
        for p, f in zip(parts, fields_per_rank):
            a = np.asarray(f[name])
            if kind == "M":
                g[..., p.own_mems] = a
            elif kind == "C":
                g[..., p.own_cells] = a[..., :p.Co]
            else:
                loc = slice((p.a - p.row_lo) * nx, (p.b - p.row_lo) * nx)
                g[..., p.a * nx:p.b * nx] = a[..., loc]
        out[name] = g
    return out
