"""Synthetic tissues at sizes the reference's seed phase cannot reach (its dense operators
need 48*C^2 bytes, SURVEY headline 3).

``make_tissue`` produces the same three groups a golden capture holds —

* ``mesh``   the ``Cells`` attribute set the loop consumes (cells.py, SURVEY §2 ★data),
* ``params`` the ``Parameters`` scalars,
* ``state``  the ``Simulator`` attributes created by ``init_core`` / ``init_dynamics``
             (sim.py:452-1012) for a fresh INIT phase,

for a jittered hexagonal sheet of cells (6 membranes each; every interior membrane has a
gap-junction partner) embedded in a square environmental grid with ~1 cell per square
(BASELINE.json configs 2-5).  The mesh is vectorised NumPy: 1 M cells take a few seconds.
Ragged cells (3-7 membranes) are exercised by the reference-built golden meshes instead.
"""
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

# neighbour k of a cell in an "odd-r" offset hex layout, counter-clockwise from east
_NB_EVEN = [(1, 0), (0, 1), (-1, 1), (-1, 0), (-1, -1), (0, -1)]
_NB_ODD = [(1, 0), (1, 1), (0, 1), (-1, 0), (0, -1), (1, -1)]


def load_profile(name):
    with open(os.path.join(_HERE, "data", "profiles.json")) as f:
        return json.load(f)[name]


def make_mesh(ncx, ncy, p, disorder=0.4, seed=20241017, grid_size=None, margin_cells=3.0):
    """Jittered hex sheet of ncx*ncy cells + env grid.  Returns the mesh dict."""
    rng = np.random.RandomState(seed)
    r = float(p["cell_radius"])
    d = 2.0 * r
    h = float(p["cell_height"])
    # lattice incl. one ring of virtual cells so that every real cell has 6 closed corners
    NX, NY = ncx + 2, ncy + 2
    ix, iy = np.meshgrid(np.arange(NX), np.arange(NY))
    px = (ix + 0.5 * (iy % 2)) * d
    py = iy * d * np.sqrt(3.0) / 2.0
    jit = 0.25 * disorder * d
    px = px + jit * (rng.rand(NY, NX) - 0.5)
    py = py + jit * (rng.rand(NY, NX) - 0.5)

    ry, rx = np.meshgrid(np.arange(1, ncy + 1), np.arange(1, ncx + 1), indexing="ij")
    ry, rx = ry.ravel(), rx.ravel()                 # real cells, row-major
    C = ncx * ncy
    cid = -np.ones((NY, NX), dtype=np.int64)
    cid[ry, rx] = np.arange(C)
    cx, cy = px[ry, rx], py[ry, rx]

    nbx = np.empty((C, 6), dtype=np.int64)
    nby = np.empty((C, 6), dtype=np.int64)
    odd = (ry % 2) == 1
    for k in range(6):
        dxe, dye = _NB_EVEN[k]
        dxo, dyo = _NB_ODD[k]
        nbx[:, k] = rx + np.where(odd, dxo, dxe)
        nby[:, k] = ry + np.where(odd, dyo, dye)
    nx_, ny_ = px[nby, nbx], py[nby, nbx]           # neighbour centres [C,6]
    nb_id = cid[nby, nbx]                           # -1 where the neighbour is virtual

    # corner k = centroid of (cell, nb_k, nb_{k+1}); membrane k = edge corner_{k-1} -> corner_k
    vx = (cx[:, None] + nx_ + np.roll(nx_, -1, axis=1)) / 3.0
    vy = (cy[:, None] + ny_ + np.roll(ny_, -1, axis=1)) / 3.0
    ax, ay = np.roll(vx, 1, axis=1), np.roll(vy, 1, axis=1)
    ex, ey = vx - ax, vy - ay
    length = np.sqrt(ex ** 2 + ey ** 2)
    mx, my = 0.5 * (vx + ax), 0.5 * (vy + ay)
    nxn, nyn = ey / length, -ex / length            # outward for counter-clockwise corners
    flip = (nxn * (mx - cx[:, None]) + nyn * (my - cy[:, None])) < 0
    nxn = np.where(flip, -nxn, nxn)
    nyn = np.where(flip, -nyn, nyn)

    M = 6 * C
    mem_to_cells = np.repeat(np.arange(C), 6)
    mem_sa = (length * h).ravel()
    R_rads = np.sqrt((mx - cx[:, None]) ** 2 + (my - cy[:, None]) ** 2).ravel()
    mem_vol = 0.5 * R_rads * mem_sa                                   # cells.py:1382
    cell_vol = mem_vol.reshape(C, 6).sum(axis=1)                      # cells.py:1385
    cell_sa = mem_sa.reshape(C, 6).sum(axis=1)
    kk = np.arange(6)[None, :]
    partner = nb_id * 6 + (kk + 3) % 6
    self_idx = np.arange(M).reshape(C, 6)
    bnd = nb_id < 0
    nn_i = np.where(bnd, self_idx, partner).ravel()
    bflags_mems = np.nonzero(bnd.ravel())[0]
    bflags_cells = np.nonzero(bnd.any(axis=1))[0]

    # ---- environmental grid (cells.py:1607-1617, finitediff.py:158-250)
    mrg = margin_cells * d
    xmin, xmax = vx.min() - mrg, vx.max() + mrg
    ymin, ymax = vy.min() - mrg, vy.max() + mrg
    side = max(xmax - xmin, ymax - ymin)
    xmax, ymax = xmin + side, ymin + side
    if grid_size is None:
        grid_size = int(np.ceil(np.sqrt(C) * side / max(vx.max() - vx.min(), vy.max() - vy.min())))
    delta = (xmax - xmin) / grid_size
    gnx = int((xmax - xmin) / delta)
    gny = int((ymax - ymin) / delta)
    xv = np.linspace(xmin, xmax, gnx + 1)
    yv = np.linspace(ymin, ymax, gny + 1)
    xc, yc = 0.5 * (xv[:-1] + xv[1:]), 0.5 * (yv[:-1] + yv[1:])

    def nearest(x, y):
        jx = np.clip(np.floor((x - xmin) / delta).astype(np.int64), 0, gnx - 1)
        jy = np.clip(np.floor((y - ymin) / delta).astype(np.int64), 0, gny - 1)
        return jy * gnx + jx
    map_mem2ecm = nearest(mx.ravel(), my.ravel())
    map_cell2ecm = nearest(cx, cy)
    E = gnx * gny
    memSa_per_envSquare = np.bincount(map_mem2ecm, weights=mem_sa, minlength=E)
    envInds_inClust = np.nonzero(np.bincount(map_mem2ecm, minlength=E))[0]
    # tight-junction index sets (cells.py:1818-1833)
    nb_of_b = nb_id[bflags_cells].ravel()
    nb_of_b = nb_of_b[nb_of_b >= 0]
    all_bound_mem_inds = map_mem2ecm.reshape(C, 6)[bflags_cells].ravel()
    interior_bound_mem_inds = map_mem2ecm.reshape(C, 6)[nb_of_b].ravel()
    ecm_inds_bound_cell = map_cell2ecm[bflags_cells]

    Xg, Yg = np.meshgrid(xc, yc)
    return {
        "mem_to_cells": mem_to_cells, "cell_mem_ptr": np.arange(0, M + 1, 6),
        "nn_i": nn_i, "bflags_mems": bflags_mems, "bflags_cells": bflags_cells,
        "map_mem2ecm": map_mem2ecm, "map_cell2ecm": map_cell2ecm,
        "mem_sa": mem_sa, "mem_nx": nxn.ravel(), "mem_ny": nyn.ravel(), "R_rads": R_rads,
        "mem_vol": mem_vol, "cell_vol": cell_vol, "cell_sa": cell_sa,
        "diviterm": cell_vol / cell_sa, "num_mems": np.full(C, 6.0),
        "delta": np.asarray(delta), "gj_len": np.asarray(2 * float(p["tm"]) + float(p["cell_space"])),
        "ecm_vol": np.asarray(h * delta ** 2), "memSa_per_envSquare": memSa_per_envSquare,
        "gj_default_weights": np.ones(M), "grid_shape": np.array([gny, gnx]),
        "cell_centres": np.column_stack((cx, cy)), "mem_mids_flat": np.column_stack((mx.ravel(), my.ravel())),
        "xypts": np.column_stack((Xg.ravel(), Yg.ravel())),
        "envInds_inClust": envInds_inClust, "all_bound_mem_inds": all_bound_mem_inds,
        "interior_bound_mem_inds": interior_bound_mem_inds, "ecm_inds_bound_cell": ecm_inds_bound_cell,
    }


def drop_membrane_pairs(mesh, frac, seed=7):
    """A RAGGED variant of the hex sheet: a random fraction of the gap-junction membrane pairs is removed, so that cells
    keep between 3 and 6 membranes (the reference's Voronoi meshes have 3-7) — the SELL-32 cell pack then holds blocks of
    different heights with padded lanes.  Geometry per cell (surface, volume, diviterm) follows the membranes that remain."""
    rng = np.random.RandomState(seed)
    M = len(mesh["mem_sa"])
    C = len(mesh["cell_vol"])
    m2c = np.asarray(mesh["mem_to_cells"])
    nn = np.asarray(mesh["nn_i"])
    me = np.arange(M)
    rep = (nn > me)                                       # one representative per interior pair
    drop = rep & (rng.rand(M) < frac)
    for _ in range(2):                                    # a cell may lose at most 3 of its 6 membranes
        both = drop | np.zeros(M, dtype=bool)
        both[nn[drop]] = True
        lost = np.bincount(m2c[both], minlength=C)
        bad = lost > 3
        drop &= ~(bad[m2c] | bad[m2c[nn]])
    gone = drop.copy()
    gone[nn[drop]] = True
    keep = ~gone
    new_index = np.cumsum(keep) - 1
    out = dict(mesh)
    for k in ("mem_to_cells", "map_mem2ecm", "mem_sa", "mem_nx", "mem_ny", "R_rads", "mem_vol", "gj_default_weights"):
        out[k] = np.asarray(mesh[k])[keep]
    out["mem_mids_flat"] = np.asarray(mesh["mem_mids_flat"])[keep]
    out["nn_i"] = new_index[nn[keep]]
    cnt = np.bincount(out["mem_to_cells"], minlength=C)
    out["cell_mem_ptr"] = np.concatenate(([0], np.cumsum(cnt)))
    out["num_mems"] = cnt.astype(float)
    out["cell_sa"] = np.bincount(out["mem_to_cells"], weights=out["mem_sa"], minlength=C)
    out["cell_vol"] = np.bincount(out["mem_to_cells"], weights=out["mem_vol"], minlength=C)
    out["diviterm"] = out["cell_vol"] / out["cell_sa"]
    Mn = len(out["mem_sa"])
    bnd = out["nn_i"] == np.arange(Mn)
    out["bflags_mems"] = np.nonzero(bnd)[0]
    out["bflags_cells"] = np.unique(out["mem_to_cells"][bnd])
    E = int(np.prod(mesh["grid_shape"]))
    out["memSa_per_envSquare"] = np.bincount(out["map_mem2ecm"], weights=out["mem_sa"], minlength=E)
    out["envInds_inClust"] = np.nonzero(np.bincount(out["map_mem2ecm"], minlength=E))[0]
    starts = out["cell_mem_ptr"][:-1]
    bc = out["bflags_cells"]
    sel = np.concatenate([np.arange(out["cell_mem_ptr"][c], out["cell_mem_ptr"][c + 1]) for c in bc]) if len(bc) < 200000 else \
        np.nonzero(np.isin(out["mem_to_cells"], bc))[0]
    out["all_bound_mem_inds"] = out["map_mem2ecm"][sel]
    return out


def make_state(mesh, p, prof):
    """Simulator attributes after init_core + init_dynamics for a fresh INIT phase
    (sim.py:452-755, 758-1012), before the first update_V."""
    I = len(prof["ions"])
    C = len(mesh["cell_vol"])
    M = len(mesh["mem_sa"])
    gny, gnx = (int(x) for x in mesh["grid_shape"])
    E = gny * gnx
    ecm = bool(p["is_ecm"])
    zs = np.asarray(prof["zs"], dtype=float)
    D_free = np.asarray(prof["D_free"], dtype=float)
    cc_cells = np.repeat(np.asarray(prof["cell_concs"], dtype=float)[:, None], C, axis=1)
    edl = E if ecm else M
    cc_env = np.repeat(np.asarray(prof["env_concs"], dtype=float)[:, None], edl, axis=1)
    st = {
        "zs": zs, "D_free": D_free, "D_gj": D_free.copy(),
        "cc_cells": cc_cells, "cc_at_mem": cc_cells[:, mesh["mem_to_cells"]],
        "cc_env": cc_env,
        "Dm_cells": np.repeat(np.asarray(prof["Dm_base"], dtype=float)[:, None], M, axis=1),
        "c_env_bound": np.asarray(prof["env_concs"], dtype=float),
        "vm": np.zeros(M), "gjopen": np.ones(M) * mesh["gj_default_weights"],
        "T": np.asarray(float(p["T"])), "rho_pump": np.asarray(1.0), "rho_channel": np.asarray(1.0),
        "NaKATP_block": np.asarray(1.0), "gj_block": np.asarray(1.0),
        "Phi_b": np.zeros(E), "vgj": np.zeros(M),
        "extra_rho_cells": np.zeros(C), "extra_J_mem": np.zeros(M), "extra_rho_env": np.zeros(edl),
        "extra_Jenv_x": np.zeros(edl), "extra_Jenv_y": np.zeros(edl),
        "bound_V": np.zeros(4), "Jn": np.zeros(M), "sigma": np.asarray(0.0),
    }
    nm = mesh["num_mems"][mesh["mem_to_cells"]]
    nfrac = float(p["smooth_cells"])
    st["smooth_weight_mem"] = (nfrac * nm - 1) / (nfrac * nm)          # sim.py:782-786
    st["smooth_weight_o"] = 1 / (nfrac * nm)
    # sim.py:974: inverse Debye length from the environmental concentrations
    st["ko_env"] = np.asarray((np.sqrt(np.dot((p["NAv"] * (p["q"] ** 2) * zs ** 2)
                                             / (p["er"] * p["eo"] * p["kb"] * p["T"]), cc_env))).mean())
    if ecm:
        # initDenv, sim.py:2353-2373
        D_env = np.empty((I, E))
        for i in range(I):
            Do = np.ones(E) * D_free[i]
            Do[mesh["envInds_inClust"]] = D_free[i] * p["D_adh"]
            tj = D_free[i] * p["D_tj"] * 1.0          # Dtj_rel = 1 (default config)
            Do[mesh["all_bound_mem_inds"]] = tj
            Do[mesh["interior_bound_mem_inds"]] = tj
            Do[mesh["ecm_inds_bound_cell"]] = tj
            D_env[i] = Do
        st["D_env"] = D_env
        st["TJ_modulator"] = np.ones((I, E))
        st["E_env_x"] = np.zeros((gny, gnx))
        st["E_env_y"] = np.zeros((gny, gnx))
    else:
        st["D_env_weight"] = np.ones((gny, gnx))
    return st


def make_tissue(n_cells, profile="mammal", ecm=True, dt=1.0e-4, seed=20241017, disorder=0.4,
                overrides=None, ragged=0.0):
    """(mesh, params, state) for a ~n_cells-cell sheet.  ``params`` are the shipped defaults of
    the reference (betse_b200/data/profiles.json) with ``dt`` and ``is_ecm`` set."""
    prof = load_profile(profile)
    p = dict(prof["p"])
    p["ions"] = np.array(prof["ions"])
    p["is_ecm"] = int(bool(ecm))
    p["dt"] = float(dt)
    p.update(overrides or {})
    ncx = int(round(np.sqrt(n_cells / (np.sqrt(3.0) / 2.0)) * (np.sqrt(3.0) / 2.0)))
    ncx = max(2, ncx)
    ncy = max(2, int(round(n_cells / ncx)))
    mesh = make_mesh(ncx, ncy, p, disorder=disorder, seed=seed)
    if ragged > 0.0:
        mesh = drop_membrane_pairs(mesh, ragged)
    # p.vol_env (no-ECM bath volume) follows the world size in the reference; keep the default
    state = make_state(mesh, p, prof)
    return mesh, p, state


# ---------------------------------------------------------------------------- BASELINE workloads beyond the ion path
BASELINE_CHANNELS = (("Nav", "Nav1p3", 2.0e-14), ("Kv", "Kv1p5", 1.0e-15), ("K_Leak", "KLeak", 0.6e-17),
                     ("Cav", "Cav1p2", 1.0e-15))


def baseline_channels(vm):
    """BASELINE configs[2]: the default general network's Nav1p3 / Kv1p5 / KLeak (sim_config.yaml) plus one voltage-gated
    Ca channel (Cav1p2, vg_ca.py:292-340), each at its steady state for the Vmem given (vg_na.py:210-228)."""
    from . import channels as chlib
    specs = []
    for name, model, dm in BASELINE_CHANNELS:
        m0, h0 = chlib.initial_state(model, vm)
        specs.append(chlib.make_channel(name, model, dm, m=m0, h=h0))
    return specs


def retarget_network(desc, C, M, E, rng, targets_every=7):
    """A recorded network description (strings, constants and tables of a reference run on a small mesh,
    betse_b200.network.unflatten) re-targeted to a synthetic tissue: same rate laws, per-cell / per-membrane /
    per-env-square tables re-drawn at the new sizes."""
    d = dict(desc)
    K = len(d["species"])
    d["c_cells"] = rng.uniform(0.05, 1.0, (K, C))
    d["growth_targets"] = [np.arange(C) for _ in range(K)]
    d["static"] = {k: (v if np.ndim(v) == 0 else np.ones(M if "mdl" in k else C)) for k, v in d["static"].items()}
    # charged substances add F*c*z to the charge (networks.py:2945): keep them dilute, the reference balances that charge at
    # set-up (networks.py:3905-3946) and a random field would not
    dilute = np.where(np.asarray(d["z"]) != 0.0, 1.0e-3, 1.0)[:, None]
    d["c_cells"] = d["c_cells"] * dilute
    if "env_on" in d:
        d["c_env"] = np.where(np.asarray(d["env_on"], dtype=bool)[:, None], rng.uniform(0.05, 0.6, (K, E)), 0.0) * dilute
        d["c_bound"] = np.asarray(d["c_bound"]) * dilute[:, 0]
        Do = np.where(np.asarray(d["D_env"]).max(axis=1) > 0, np.asarray(d["D_env"]).max(axis=1), 0.0)
        d["D_env"] = Do[:, None] * rng.uniform(0.2, 1.0, (K, E))
    if "c_mems" in d:
        d["c_mems"] = rng.uniform(0.05, 1.0, (K, M)) * dilute
    if "transporters" in d:
        d["transporters"] = [dict(t, targets_cell=np.arange(0, C, targets_every if j else 1), targets_mem=np.arange(M),
                                  targets_env=np.arange(E)) for j, t in enumerate(d["transporters"])]
    return d
