"""TissueEngine — Python owner of one ``betse_ctx`` (one GPU, one simulation).

It speaks the reference's vocabulary: the mesh is the set of ``Cells`` attributes the loop
consumes, parameters are the ``Parameters`` scalars, state arrays carry the names of the
``Simulator`` attributes (``cc_cells``, ``cc_at_mem``, ``cc_env``, ``vm``, ``gjopen``,
``Dm_cells``, ``D_env``, ``TJ_modulator``, ``E_env_x`` ...).  All compute happens in
libbetse_b200.so; nothing here falls back to NumPy.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import BetseB200Error

# Simulator attribute -> (StateHost member, shape key)
_DOWN_SHAPES = {
    "cc_cells": "IC", "cc_env": "IE", "vm": "M", "gjopen": "M", "Dm_cells": "IM",
    "E_env_x": "E", "E_env_y": "E", "v_env": "E", "rho_env": "E", "rho_cells": "C", "vm_ave": "C",
    "fluxes_mem": "IM", "fluxes_gj": "IM", "fluxes_env_x": "IE", "fluxes_env_y": "IE",
    "rate_NaKATP": "M", "Jmem": "M", "Jgj": "M", "Jn": "M", "I_mem": "M", "Jc": "M", "Emc": "M",
    "dvm": "M", "J_cell_x": "C", "J_cell_y": "C", "E_cell_x": "C", "E_cell_y": "C",
    "sigma_cell": "C", "E_gj_x": "M", "E_gj_y": "M",
    "J_env_x": "E", "J_env_y": "E", "B_field": "E", "Jtx": "E", "Jty": "E", "Phi_b": "E",
    "extra_rho_cells": "C", "extra_rho_env": "E", "extra_J_mem": "M", "D_env_eff": "IE",
}
# what get_current leaves on the env grid of a tissue WITHOUT extracellular spaces (ion_current.py:116-158)
_NOECM_FIELD = ("v_env", "E_env_x", "E_env_y", "J_env_x", "J_env_y", "B_field", "Jtx", "Jty")
DIAG_FIELDS = ("fluxes_mem", "fluxes_gj", "fluxes_env_x", "fluxes_env_y", "rate_NaKATP", "Jmem",
               "Jgj", "Jn", "I_mem", "Jc", "Emc", "dvm", "J_cell_x", "J_cell_y", "E_cell_x",
               "E_cell_y", "sigma_cell", "vm_ave", "E_gj_x", "E_gj_y", "J_env_x", "J_env_y", "B_field", "Jtx", "Jty")


def trs_needs_vol(net):
    """Transporters upload mem_sa/mem_vol themselves (set_network)."""
    return bool(net.get("transporters"))


def gaussian_taps():
    """Taps of scipy.ndimage.gaussian_filter(sigma=1) (truncate=4 -> radius 4), formed exactly
    like scipy's _gaussian_kernel1d so that v_env matches ion_current.py:104."""
    x = np.arange(-4, 5)
    phi = np.exp(-0.5 / 1.0 * x ** 2)
    phi = phi / phi.sum()
    return phi[4:].copy()


def _scalar(v):
    return float(np.asarray(v).reshape(-1)[0]) if np.ndim(v) else float(v)


def _pinned_array(lib, ptr, n, shape):
    """NumPy view of page-locked memory that frees it (betse_host_free) when the last view dies — an array lent to the
    Simulator at the end of a phase simply outlives the engine, no copy."""
    import weakref
    raw = (C.c_double * n).from_address(ptr.value)
    fin = weakref.finalize(raw, lib.betse_host_free, C.c_void_p(ptr.value))
    fin.atexit = False              # at interpreter exit the OS reclaims it; the CUDA runtime may already be gone
    return np.ctypeslib.as_array(raw).reshape(shape)


class PinnedPrefetch:
    """Page-locked staging for the sampled-step downloads, allocated on a helper thread: pinning a few hundred MB costs
    as much as uploading the tissue, so the loop starts it before it builds the engine (TissueEngine.adopt_pinned)."""

    def __init__(self, device, shapes):
        import threading
        self.lib = capi.load()
        self.device, self.shapes, self.got = int(device), dict(shapes), {}
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        for name, shape in self.shapes.items():
            n = int(np.prod(shape))
            ptr = C.c_void_p()
            if self.lib.betse_host_alloc_on(self.device, C.c_size_t(n * 8), C.byref(ptr)) != 0 or not ptr.value:
                return                  # the engine allocates what is missing (and reports the failure) itself
            self.got[name] = _pinned_array(self.lib, ptr, n, shape)

    def join(self):
        self.thread.join()
        got, self.got = self.got, {}
        return got

    def release(self):
        self.join().clear()


class TissueEngine:
    def __init__(self, mesh, params, state=None, device=0, partition=None):
        """``mesh``: dict of Cells arrays (keys as in oracle/refrun.py CELLS_FIELDS);
        ``params``: dict of Parameters scalars + ``ions``; ``state``: dict of Simulator
        attributes at loop entry (optional, may be uploaded later)."""
        self.lib = capi.load()
        if self.lib.betse_device_count() <= 0:
            raise BetseB200Error("no CUDA device: betse_b200 has no CPU fallback")
        self._keep = []
        self.mesh = mesh
        self.p = {k: (np.asarray(v).item() if np.asarray(v).ndim == 0 and np.asarray(v).dtype.kind != "U"
                      else np.asarray(v)) for k, v in params.items()}
        self.is_ecm = bool(self.p["is_ecm"])
        self.mem_to_cells = capi.as_i32(mesh["mem_to_cells"])
        self.cell_mem_ptr = capi.as_i32(mesh["cell_mem_ptr"])
        self.Co = len(self.cell_mem_ptr) - 1                      # owned cells
        self.C = int((partition or {}).get("n_cells", self.Co))   # + ghost cells (domain decomposition)
        self.M = len(self.mem_to_cells)
        gs = np.asarray(mesh["grid_shape"]).astype(int)
        self.ny, self.nx = int(gs[0]), int(gs[1])
        self.E = self.ny * self.nx
        ions = [str(x) for x in self.p["ions"]]
        self.ions = ions
        self.I = len(ions)
        self._sched = {}
        st = state or {}
        self._hp = self._make_params(st)
        m = self._make_mesh(partition)
        ctx = C.c_void_p()
        rc = self.lib.betse_create(C.byref(ctx), C.byref(m), C.byref(self._hp), int(device))
        if rc != 0:
            buf = C.create_string_buffer(1024)
            self.lib.betse_create_error(buf, 1024)
            raise BetseB200Error("betse_create failed (%d): %s" % (rc, buf.value.decode()))
        self.ctx = ctx
        self.steps_done = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        if state:
            self.upload(**state)

    # ------------------------------------------------------------------ construction helpers
    def _make_mesh(self, partition):
        mesh, p = self.mesh, self.p
        m = capi.Mesh()
        k = self._keep

        def f64(name, required=True):
            if name not in mesh:
                if required:
                    raise KeyError("mesh is missing %r" % name)
                return None
            a = capi.as_f64(mesh[name])
            k.append(a)
            return capi.ptr_f64(a)

        def i32(a):
            a = capi.as_i32(a)
            k.append(a)
            return capi.ptr_i32(a)
        m.n_cells, m.n_mems, m.ny, m.nx = self.C, self.M, self.ny, self.nx
        m.mem_to_cells = i32(self.mem_to_cells)
        m.cell_mem_ptr = i32(self.cell_mem_ptr)
        # the reference's index arrays are int64: narrowing them (and the mean below) are memory-bound passes over M
        # entries, run side by side (NumPy releases the GIL inside them)
        from concurrent.futures import ThreadPoolExecutor

        def memsa_mean():
            if "memsa_mean" in mesh or "memSa_per_envSquare" not in mesh:
                return None
            msa = np.asarray(mesh["memSa_per_envSquare"], dtype=np.float64)
            return float(msa[np.asarray(mesh["map_mem2ecm"])].mean())
        with ThreadPoolExecutor(3) as pool:
            f_nn = pool.submit(capi.as_i32, mesh["nn_i"])
            f_m2e = pool.submit(capi.as_i32, mesh["map_mem2ecm"])
            f_mean = pool.submit(memsa_mean)
            bf = np.zeros(self.M, dtype=np.uint8)
            bfl = np.asarray(mesh["bflags_mems"])
            if bfl.dtype == np.bool_ and bfl.size == self.M:
                bf[:] = bfl
            else:
                bf[bfl.astype(np.int64)] = 1       # the reference stores a list of membrane indices
            m.nn_i = i32(f_nn.result())
            m.map_mem2ecm = i32(f_m2e.result())
            mean = f_mean.result()
        k.append(bf)
        m.bflags_mems = bf.ctypes.data_as(C.POINTER(C.c_uint8))
        m.mem_sa, m.mem_nx, m.mem_ny = f64("mem_sa"), f64("mem_nx"), f64("mem_ny")
        m.R_rads = f64("R_rads", required=False)
        m.cell_vol, m.cell_sa, m.diviterm = f64("cell_vol"), f64("cell_sa"), f64("diviterm")
        nm = capi.as_f64(np.asarray(mesh["num_mems"], dtype=np.float64))
        k.append(nm)
        m.num_mems = capi.ptr_f64(nm)
        m.memSa_per_envSquare = f64("memSa_per_envSquare", required=False)
        m.gj_default_weights = f64("gj_default_weights", required=False)
        m.delta = float(mesh["delta"])
        m.gj_len = float(mesh["gj_len"])
        m.ecm_vol = float(mesh["ecm_vol"]) if "ecm_vol" in mesh else float(p["cell_height"]) * m.delta ** 2
        if "memsa_mean" in mesh:          # a strip of a decomposed tissue carries the global mean
            m.memsa_mean = float(mesh["memsa_mean"])
        elif mean is not None:
            m.memsa_mean = mean
        else:
            m.memsa_mean = 1.0
        part = partition or {}
        m.n_cells_owned = int(part.get("n_cells_owned", self.C))
        m.n_mems_owned = int(part.get("n_mems_owned", self.M))
        m.n_flux_slots = int(part.get("n_flux_slots", self.M))
        m.y0 = int(part.get("y0", 0))
        m.ny_global = int(part.get("ny_global", self.ny))
        m.y_own0 = int(part.get("y_own0", 0))
        m.y_own1 = int(part.get("y_own1", self.ny))
        if "ecm_slot_ptr" in part:
            m.ecm_slot_ptr = i32(part["ecm_slot_ptr"])
            m.ecm_slot_idx = i32(part["ecm_slot_idx"])
        return m

    def _make_params(self, st):
        p = self.p
        hp = capi.Params()
        hp.abi_version = capi.ABI_VERSION
        hp.n_ions = self.I
        idx = {n: i for i, n in enumerate(self.ions)}
        hp.iNa, hp.iK = idx.get("Na", -1), idx.get("K", -1)
        hp.iCa, hp.iP = idx.get("Ca", -1), idx.get("P", -1)
        S = self._sched
        if "zs" in st:
            S["zs"] = np.asarray(st["zs"], dtype=float)
            S["D_free"] = np.asarray(st["D_free"], dtype=float)
            dgj = np.asarray(st["D_gj"], dtype=float)
            if dgj.ndim == 2:
                if not np.all(dgj == dgj[:, :1]):
                    raise BetseB200Error("non-uniform sim.D_gj is not supported")
                dgj = dgj[:, 0]
            S["D_gj"] = dgj
        for key, default in (("c_env_bound", np.zeros(self.I)), ("T", p.get("T", 310.0)),
                             ("ko_env", 1.0), ("rho_pump", 1.0), ("rho_channel", 1.0),
                             ("bound_V", np.zeros(4)), ("sigma", 0.0)):
            if key in st:
                S[key] = np.asarray(st[key], dtype=float)
            S.setdefault(key, np.asarray(default, dtype=float))
        for key in ("NaKATP_block", "gj_block"):
            if key in st and np.ndim(st[key]) == 0:
                S[key] = float(st[key])
            S.setdefault(key, 1.0)
        if "zs" not in S:
            raise BetseB200Error("state must provide zs, D_free and D_gj at construction")
        for i in range(self.I):
            hp.z[i] = S["zs"][i]
            hp.D_free[i] = S["D_free"][i]
            hp.D_gj[i] = S["D_gj"][i]
            hp.c_env_bound[i] = np.asarray(S["c_env_bound"]).reshape(-1)[i]
        if not self.is_ecm and "cc_env" in st:
            ce = np.asarray(st["cc_env"], dtype=float)
            for i in range(self.I):
                hp.cenv_uniform[i] = ce[i].reshape(-1)[0]
        for f in ("F", "R", "q", "kb", "eo", "er", "cm", "tm", "NAv", "mu", "dt", "alpha_NaK",
                  "alpha_Ca", "KmNK_Na", "KmNK_K", "KmNK_ATP", "KmCa_Ca", "KmCa_ATP", "cATP", "cADP",
                  "cPi", "deltaGATP", "gj_surface", "gj_vthresh", "gj_min", "cell_height", "vol_env",
                  "cell_radius", "true_cell_size", "sharpness", "cell_polarizability", "smooth_cells"):
            setattr(hp, f, float(p[f]))
        hp.T_sim = _scalar(S["T"])
        hp.T_p = float(p["T"])
        hp.rho_pump, hp.rho_channel = _scalar(S["rho_pump"]), _scalar(S["rho_channel"])
        hp.ko_env = _scalar(S["ko_env"])
        hp.sigma_env = _scalar(S["sigma"])
        bv = np.asarray(S["bound_V"], dtype=float).reshape(-1)
        for i in range(4):
            hp.bound_V[i] = bv[i]
        gw = gaussian_taps()
        for i in range(5):
            hp.gauss_w[i] = gw[i]
        hp.NaKATP_block_scalar = float(S["NaKATP_block"])
        hp.gj_block_scalar = float(S["gj_block"])
        hp.is_ecm = int(bool(p["is_ecm"]))
        hp.v_sensitive_gj = int(bool(p["v_sensitive_gj"]))
        hp.cluster_open = int(bool(p.get("cluster_open", 1)))
        hp.fast_update_ecm = int(bool(p.get("fast_update_ecm", 0)))
        return hp

    def _check(self, rc, what):
        if rc != 0:
            buf = C.create_string_buffer(1024)
            self.lib.betse_last_error(self.ctx, buf, 1024)
            raise BetseB200Error("%s failed (%d): %s" % (what, rc, buf.value.decode()))

    # ------------------------------------------------------------------ state movement
    def upload(self, **state):
        """Upload Simulator attributes (any subset).  Unknown names are ignored so that a whole
        capture group can be passed."""
        sh = capi.StateHost()
        keep = []

        def put(member, arr, n):
            a = capi.as_f64(arr).reshape(-1)
            if a.size != n:
                raise BetseB200Error("%s: expected %d values, got %d" % (member, n, a.size))
            keep.append(a)
            self.h2d_bytes += a.nbytes
            setattr(sh, member, capi.ptr_f64(a))
        I, Cn, M, E = self.I, self.C, self.M, self.E
        starts = self.cell_mem_ptr[:-1]
        if "cc_cells" in state:
            put("cc_cells", state["cc_cells"], I * Cn)
        if "cc_mid" in state:           # per-cell form (incl. ghost cells), partition.py
            put("cc_at_mem_cell", state["cc_mid"], I * Cn)
        elif "cc_at_mem" in state:
            cam = np.asarray(state["cc_at_mem"], dtype=float)
            put("cc_at_mem_cell", cam[:, starts] if cam.shape[1] == M else cam, I * Cn)
        if self.is_ecm:
            if "cc_env" in state:
                put("cc_env", state["cc_env"], I * E)
            if "D_env" in state:
                d = np.asarray(state["D_env"], dtype=float).reshape(I, E)
                tj = np.asarray(state.get("TJ_modulator", getattr(self, "_tj", 1.0)), dtype=float)
                self._denv = d
                self._tj = tj
                put("D_env_eff", d * tj.reshape(d.shape) if np.ndim(tj) else d * tj, I * E)
            elif "TJ_modulator" in state:
                self._tj = np.asarray(state["TJ_modulator"], dtype=float)
                put("D_env_eff", self._denv * self._tj.reshape(self._denv.shape), I * E)
            if "E_env_x" in state:
                put("E_env_x", state["E_env_x"], E)
                put("E_env_y", state["E_env_y"], E)
        elif "cc_env" in state:
            ce = np.asarray(state["cc_env"], dtype=float)
            put("cenv_uniform", ce.reshape(I, -1)[:, 0], I)
        if "E_cell_x" in state and "E_cell_y" in state and "mem_nx" in self.mesh:
            # sim.Emc as update_V left it (sim.py:2079-2080): what update_intra of a charged substance reads in the
            # first step (set_network hands it on)
            ex, ey = np.asarray(state["E_cell_x"], dtype=float), np.asarray(state["E_cell_y"], dtype=float)
            if ex.shape == (self.Co,) and ey.shape == (self.Co,):
                m2c = self.mem_to_cells.astype(np.int64)
                self._Emc0 = ex[m2c] * np.asarray(self.mesh["mem_nx"], dtype=float) + ey[m2c] * np.asarray(self.mesh["mem_ny"], dtype=float)
        if "Phi_b" in state:
            put("Phi_b", state["Phi_b"], E)
        if "D_env_weight" in state and not self.is_ecm:     # no-ECM field diagnostics (ion_current.py:116-158)
            put("D_env_weight", state["D_env_weight"], E)
            self.noecm_field = True
        if "vm_cell" in state:
            put("vm_cell", state["vm_cell"], Cn)
        elif "vm" in state:
            put("vm", state["vm"], M)
        if "gjopen" in state:
            put("gjopen", state["gjopen"], M)
        if "Dm_cells" in state:
            put("Dm_cells", state["Dm_cells"], I * M)
        for name, n in (("extra_rho_cells", Cn), ("extra_rho_env", E), ("extra_J_mem", M)):
            if name in state and np.any(np.asarray(state[name]) != 0):
                put(name, state[name], n)
        resched = False
        blk_arrays = self.__dict__.setdefault("_block_arrays", set())
        for name in ("NaKATP_block", "gj_block"):
            if name in state:
                if np.ndim(state[name]) == 0 or np.size(state[name]) == 1:
                    v = _scalar(state[name])
                    if name in blk_arrays:
                        # the device holds a per-membrane array for this block (sim.NaKATP_block starts as np.ones(mdl),
                        # sim.py:848-849) and the kernels read the array once it exists: a scalar rebinding of the
                        # Simulator attribute (tishandler.py:790) replaces it as a whole
                        put(name, np.full(M, v), M)
                    elif v != self._sched.get(name):
                        self._sched[name] = v
                        resched = True
                else:
                    put(name, state[name], M)
                    blk_arrays.add(name)
        for name in ("c_env_bound", "T", "bound_V", "D_gj"):
            if name in state:
                resched |= self._set_sched_value(name, state[name])
        self._check(self.lib.betse_upload_state(self.ctx, C.byref(sh)), "betse_upload_state")
        if resched:
            self._push_schedule()

    def _set_sched_value(self, name, value):
        v = np.asarray(value, dtype=float)
        if name == "D_gj" and v.ndim == 2:
            if not np.all(v == v[:, :1]):
                raise BetseB200Error("non-uniform sim.D_gj is not supported")
            v = v[:, 0]
        old = self._sched.get(name)
        if old is not None and np.shape(old) == v.shape and np.array_equal(old, v):
            return False
        self._sched[name] = v.copy()
        return True

    def _push_schedule(self):
        self._hp = self._make_params({})
        self._check(self.lib.betse_set_schedule(self.ctx, C.byref(self._hp)), "betse_set_schedule")

    def set_field(self, name, value):
        """Apply one of the quantities TissueHandler.fire_events rewrites."""
        if name in ("c_env_bound", "T", "bound_V", "D_gj"):
            if self._set_sched_value(name, value):
                self._push_schedule()
        else:
            self.upload(**{name: value})

    def set_bound_V(self, v):
        self.set_field("bound_V", v)

    def set_bath(self, values):
        """Tissue WITHOUT extracellular spaces: overwrite the well-mixed bath concentration of some ions
        (``{ion index: value}``) — the global K_env / Cl_env / Na_env events write ``sim.cc_env[ion][:]`` on every step
        (tishandler.py:759-777); the other ions keep the value the device has advanced."""
        if self.is_ecm:
            raise BetseB200Error("set_bath applies to tissues without extracellular spaces")
        ce = np.full(self.I, np.nan)
        for i, v in values.items():
            ce[int(i)] = float(v)
        sh = capi.StateHost()
        sh.cenv_uniform = capi.ptr_f64(ce)
        self.h2d_bytes += 8 * len(values)
        self._check(self.lib.betse_upload_state(self.ctx, C.byref(sh)), "betse_upload_state")

    # ------------------------------------------------------------------ stepping
    def step(self, n=1, diag=False):
        st = C.c_uint32(0)
        self._check(self.lib.betse_step(self.ctx, int(n), capi.STEP_DIAG if diag else 0, C.byref(st)),
                    "betse_step")
        self.steps_done += n
        return int(st.value)

    def update_V(self):
        """Simulator.update_V before the loop (sim.py:1041): charge, Vmem, env field from the
        uploaded concentrations."""
        self._check(self.lib.betse_update_v(self.ctx), "betse_update_v")

    def profile(self, n):
        """Run ``n`` steps timed on the device.  Returns (total_ms, {kernel: ms_per_launch})."""
        tot = C.c_float(0)
        kms = (C.c_float * capi.NKERNELS)()
        kl = (C.c_int * capi.NKERNELS)()
        self._check(self.lib.betse_step_profile(self.ctx, int(n), C.byref(tot), C.byref(kms), C.byref(kl)),
                    "betse_step_profile")
        self.steps_done += n + min(n, 20)
        names = [self.lib.betse_kernel_name(k).decode() for k in range(capi.NKERNELS)]
        return float(tot.value), {names[k]: float(kms[k]) for k in range(capi.NKERNELS) if kl[k]}

    def adopt_pinned(self, prefetch):
        """Take over staging that a PinnedPrefetch allocated while this engine was being built."""
        self.__dict__["_prefetch"] = prefetch

    def _pinned_buffer(self, name, shape):
        """A reusable page-locked array for ``name`` (betse_host_alloc); freed by close()."""
        pool = self.__dict__.setdefault("_pinned", {})
        pf = self.__dict__.pop("_prefetch", None)
        if pf is not None:
            for k, arr in pf.join().items():
                pool.setdefault(k, arr)
        if name in pool and pool[name].shape != tuple(shape):
            del pool[name]
        if name not in pool:
            n = int(np.prod(shape))
            ptr = C.c_void_p()
            if self.lib.betse_host_alloc(C.c_size_t(n * 8), C.byref(ptr)) != 0 or not ptr.value:
                raise BetseB200Error("betse_host_alloc(%d bytes) failed" % (n * 8))
            pool[name] = _pinned_array(self.lib, ptr, n, shape)
        return pool[name]

    def download(self, fields=("cc_cells", "cc_at_mem", "cc_env", "vm", "gjopen", "rho_cells",
                               "E_env_x", "E_env_y", "v_env", "rho_env"), pinned=False):
        """Fetch Simulator attributes by name -> dict of NumPy arrays in the reference's shapes.

        ``pinned=True`` returns views of page-locked staging buffers that the NEXT pinned download of the
        same field overwrites and that die with the engine — for consumers that copy what they keep, like
        ``Simulator.write2storage`` (sim.py:1789-1884: np.copy / ``*1`` of every array it stores)."""
        sh = capi.StateHost()
        out = {}
        I, Cn, M, E = self.I, self.C, self.M, self.E
        shapes = {"IC": (I, Cn), "IE": (I, E), "IM": (I, M), "M": (M,), "C": (Cn,), "E": (E,)}
        want_cam = False
        cenv = None
        for f in fields:
            if f == "cc_at_mem":
                want_cam = True
                buf = np.empty((I, Cn))
                sh.cc_at_mem_cell = capi.ptr_f64(buf)
                out["_cam"] = buf
                continue
            if f in ("cc_env",) and not self.is_ecm:
                cenv = np.empty(I)
                sh.cenv_uniform = capi.ptr_f64(cenv)
                continue
            if not self.is_ecm and _DOWN_SHAPES.get(f, "").endswith("E") and f != "Phi_b" and not (
                    getattr(self, "noecm_field", False) and f in _NOECM_FIELD):
                continue
            if f not in _DOWN_SHAPES:
                raise KeyError(f)
            # page-locking costs ~0.5 s/GB: it only pays for staging that many samples reuse (pin_staging, set by the loop)
            buf = self._pinned_buffer(f, shapes[_DOWN_SHAPES[f]]) if (pinned and getattr(self, "pin_staging", True)) \
                else np.empty(shapes[_DOWN_SHAPES[f]])
            setattr(sh, f, capi.ptr_f64(buf))
            out[f] = buf
        self._check(self.lib.betse_download_sample(self.ctx, C.byref(sh)), "betse_download_sample")
        self.d2h_bytes += sum(a.nbytes for a in out.values()) + (cenv.nbytes if cenv is not None else 0)
        if want_cam:
            cam = out.pop("_cam")
            out["cc_at_mem"] = np.empty((I, M))
            self.lib.betse_host_expand(capi.ptr_f64(out["cc_at_mem"]), capi.ptr_f64(cam), capi.ptr_i32(self.mem_to_cells),
                                       I, Cn, M)
        if cenv is not None:
            out["cc_env"] = np.repeat(cenv[:, None], M, axis=1)   # the reference keeps [I,M] (sim.py:487-490)
        return out

    # ------------------------------------------------------------------ fast (equivalent-circuit) solver
    FAST_FIELDS = {"vm_ave": "C", "gjopen": "M", "vgj": "M", "Jn": "M", "Emx": "M", "Emy": "M",
                   "J_cell_x": "C", "J_cell_y": "C", "E_cell_x": "C", "E_cell_y": "C"}

    def fast_setup(self, state):
        """The constants Simulator.fast_sim_init left on the Simulator (sim.py:1393-1452: G_Leak, E_Leak, G_gj) and the
        state the fast loop advances (vm_ave, gjopen); optional ``extra_J_mem``."""
        sh = capi.FastHost()
        keep = []
        n_of = {"vm_ave": self.C, "gjopen": self.M, "G_Leak": self.C, "E_Leak": self.C, "G_gj": self.C,
                "sigma_cell": self.C, "extra_J_mem": self.M}
        for name, n in n_of.items():
            if name not in state or state[name] is None:
                if name == "extra_J_mem":
                    continue
                raise BetseB200Error("fast solver: sim.%s is missing (Simulator.fast_sim_init has not run?)" % name)
            a = capi.as_f64(np.broadcast_to(np.asarray(state[name], dtype=float), (n,)) if np.ndim(state[name]) == 0
                            else state[name]).reshape(-1)
            if a.size != n:
                raise BetseB200Error("fast solver: %s has %d values, expected %d" % (name, a.size, n))
            if name == "extra_J_mem" and not np.any(a):
                continue
            keep.append(a)
            self.h2d_bytes += a.nbytes
            setattr(sh, name, capi.ptr_f64(a))
        self._check(self.lib.betse_fast_setup(self.ctx, C.byref(sh)), "betse_fast_setup")

    def fast_set_channels(self, cbar, rev_E, geo_conv=1.0):
        """Channels (set_channels) under the fast solver, run_fast_loop_channels (networks.py:3217-3280): ``cbar`` / ``rev_E``
        [n_ions] = sim.cbar_dic / sim.rev_E_dic of Simulator.fast_sim_init in ion-index order.  Before fast_setup."""
        P = self.p
        z = np.array([float(self._hp.z[i]) for i in range(self.I)])
        # stb.get_conductivity(D, z, c, d, p) = (D*q*z^2*F*c)/(d*kb*T) without D, times sim.geo_conv (networks.py:3267)
        coef = capi.as_f64((float(P["q"]) * (z ** 2) * float(P["F"]) * np.asarray(cbar, dtype=float)) /
                           (float(P["tm"]) * float(P["kb"]) * float(P["T"])) * float(geo_conv))
        rev = capi.as_f64(np.asarray(rev_E, dtype=float))
        if coef.size != self.I or rev.size != self.I:
            raise BetseB200Error("fast solver channels: cbar / rev_E need one value per ion")
        self._check(self.lib.betse_fast_set_channels(self.ctx, capi.ptr_f64(coef), capi.ptr_f64(rev)), "betse_fast_set_channels")

    def fast_step(self, n=1, diag=False):
        """n iterations of Simulator._run_fast_sim_core_loop's body (sim.py:1547-1592); ``diag``: the currents and fields
        of the last one are formed as well (sampled steps)."""
        st = C.c_uint32(0)
        self._check(self.lib.betse_fast_step(self.ctx, int(n), capi.STEP_DIAG if diag else 0, C.byref(st)), "betse_fast_step")
        self.steps_done += n
        return int(st.value)

    def fast_download(self, fields=("vm_ave", "gjopen")):
        sh = capi.FastHost()
        out = {}
        for f in fields:
            n = self.C if self.FAST_FIELDS[f] == "C" else self.M
            out[f] = np.empty(n)
            setattr(sh, f, capi.ptr_f64(out[f]))
        self._check(self.lib.betse_fast_download(self.ctx, C.byref(sh)), "betse_fast_download")
        self.d2h_bytes += sum(a.nbytes for a in out.values())
        if "vm_ave" in out:
            out["vm"] = out["vm_ave"][self.mem_to_cells]          # sim.py:1563
        return out

    # ------------------------------------------------------------------ voltage-gated channels
    def set_channels(self, specs, phase_init=False, affect_charge=None):
        """``specs``: channel dicts (betse_b200.channels.make_channel / the reference's
        ``MasterOfNetworks.channels`` via simloop.channels_from_sim), applied in order.  In the
        INIT phase channels with ``init_active`` False are skipped (networks.py:3140-3141)."""
        from . import channels as chlib
        specs = [c for c in specs if not (phase_init and not c["init_active"])]
        self.channel_names = [c.get("name", "chan%d" % k) for k, c in enumerate(specs)]
        idx = {n: i for i, n in enumerate(self.ions)}
        # one C-ABI entry per (channel, conducted ion): multi-ion families (vg_funny, cation) follow their first
        # entry with `same_gates` entries (networks.py:3158-3203: the gates advance once, each ion gets its flux)
        entries, self._chan_entry = [], []
        for c in specs:
            if c["model"] not in chlib.MODELS:
                raise BetseB200Error("channel type %r is not tabulated (betse_b200/channels.py)" % c["model"])
            ions, perms = chlib.ions_of(c["model"])
            perms = list(c.get("rel_perms") or ([float(c.get("rel_perm", 1.0))] if len(ions) == 1 else perms))
            for ion in ions:
                if ion not in idx:
                    raise BetseB200Error("channel %r conducts %s, which this ion profile does not simulate"
                                         % (c.get("name"), ion))
            self._chan_entry.append(len(entries))
            entries += [(c, ion, float(pm), j) for j, (ion, pm) in enumerate(zip(ions, perms))]
        arr = (capi.Channel * max(1, len(entries)))()
        keep = []
        for k, (c, ion_name, rel_perm, j) in enumerate(entries):
            mdl = chlib.MODELS[c["model"]]
            d = arr[k]
            d.handler = int(c.get("handler", 0))
            d.mod_prog = int(c.get("mod_prog", -1))
            d.ion, d.rel_perm, d.same_gates = idx[ion_name], rel_perm, int(j > 0)
            if j > 0:
                continue
            d.mpower, d.hpower = int(mdl["mpow"]), int(mdl["hpow"])
            for q, spec in enumerate(chlib.device_quantities(c["model"])):
                d.kind[q] = chlib.KIND[spec[0]]
                terms = [spec[1], spec[2] if len(spec) > 2 else (0, 0.0, 0.0, 0.0, 0.0)]
                for dst, t in zip((d.a[q], d.b[q]), terms):
                    dst.type = int(t[0])
                    for j in range(4):
                        dst.p[j] = float(t[1 + j])
            d.time_unit, d.max_Dm = float(mdl["time_unit"]), float(c["maxDm"])
            d.v_shift = float(mdl["shift"])
            m0, h0 = np.ones(self.M), np.ones(self.M)
            tg = c.get("targets")
            if tg is None:
                tg = np.arange(self.M)
            else:
                tg = np.asarray(tg).astype(np.int64)
                mask = np.zeros(self.M, dtype=np.uint8)
                mask[tg] = 1
                keep.append(mask)
                d.target_mask = mask.ctypes.data_as(C.POINTER(C.c_uint8))
            if c.get("m") is None:      # fresh channel: steady state at the current Vmem (vg_na.py:210-228)
                vm = self.download(["vm"])["vm"][tg]
                cm, chh = chlib.initial_state(c["model"], vm)
            else:
                cm, chh = c["m"], c["h"]
            m0[tg] = np.asarray(cm, dtype=float) * np.ones(len(tg))
            h0[tg] = np.asarray(chh, dtype=float) * np.ones(len(tg))
            keep += [m0, h0]
            d.m0, d.h0 = capi.ptr_f64(m0), capi.ptr_f64(h0)
        if affect_charge is None:
            affect_charge = bool(self.p.get("substances_affect_charge", 0))
        self._check(self.lib.betse_set_channels(self.ctx, len(entries), arr, int(bool(affect_charge))), "betse_set_channels")
        self.n_channels = len(specs)

    def channel_state(self, k):
        """{'m','h','P','flux','DChan'} of channel ``k`` ([M] each)."""
        out = {f: np.empty(self.M) for f in ("m", "h", "P", "flux", "DChan")}
        k = self._chan_entry[int(k)]         # flux / DChan: the last conducted ion's, as the reference leaves them (networks.py:3201-3203)
        self._check(self.lib.betse_channel_state(self.ctx, int(k), *(capi.ptr_f64(out[f]) for f in ("m", "h", "P", "flux", "DChan"))),
                    "betse_channel_state")
        return out

    def tj_modulator(self):
        """sim.TJ_modulator [I][E] as the tight-junction modulators of a network left it (networks.py:3301-3317); None
        without such modulators."""
        if getattr(self, "_tj_targets", None) is None:
            return None
        out = np.empty((self.I, self.E))
        self._check(self.lib.betse_network_tj_modulator(self.ctx, capi.ptr_f64(out)), "betse_network_tj_modulator")
        self.d2h_bytes += out.nbytes
        return out

    # ------------------------------------------------------------------ general / gene network
    def set_network(self, net, handler=0):
        """``net``: a compiled network (betse_b200.network.compile_network): substances, rate
        programs, tables.  ``handler`` 0 = general network, 1 = gene regulatory network."""
        from . import ratelaw
        K = len(net["species"])
        programs = list(net["rate_programs"]) + list(net["mod_programs"])
        code, ptr = ratelaw.pack_programs(programs)
        tabs = net["tables"]
        n = capi.Network()
        keep = []

        def f64(a):
            a = capi.as_f64(a)
            keep.append(a)
            return capi.ptr_f64(a)
        n.n_species, n.n_rates, n.n_programs = K, len(net["rate_programs"]), len(programs)
        n.n_consts, n.n_cell_arrays, n.n_mem_arrays = len(tabs.consts), len(tabs.cell_arrays), len(tabs.mem_arrays)
        c0 = np.zeros((K, self.C))
        c0[:, :self.Co] = np.asarray(net["c_cells"], dtype=float).reshape(K, -1)
        n.c_cells = f64(c0)
        keep += [code, ptr]
        n.code, n.prog_ptr = capi.ptr_i32(code), capi.ptr_i32(ptr)
        n.consts = f64(np.asarray(tabs.consts if tabs.consts else [0.0]))
        if tabs.cell_arrays:
            n.cell_arrays = f64(np.stack(tabs.cell_arrays))
        if tabs.mem_arrays:
            n.mem_arrays = f64(np.stack(tabs.mem_arrays))
        gm = net.get("growth_mask")
        if gm is not None:
            gm = np.ascontiguousarray(gm, dtype=np.uint8).reshape(K, self.C)
            keep.append(gm)
            n.growth_mask = gm.ctypes.data_as(C.POINTER(C.c_uint8))
        n.stoich = f64(np.asarray(net["stoich"], dtype=float).reshape(K, n.n_rates))
        n.Dgj, n.z, n.time_factor = f64(net["Dgj"]), f64(net["z"]), f64(net["time_factor"])
        env_on = np.asarray(net.get("env_on", np.zeros(K)), dtype=np.uint8).reshape(K)
        if env_on.any():       # membrane / extracellular legs of molecule_mover (sim_toolbox.py:909-1153)
            if not self.is_ecm:
                raise BetseB200Error("network substances in the environment need extracellular spaces")
            eo = np.ascontiguousarray(env_on)
            keep.append(eo)
            n.env_on = eo.ctypes.data_as(C.POINTER(C.c_uint8))
            n.Dm, n.c_bound = f64(np.asarray(net["Dm"], dtype=float).reshape(K)), f64(np.asarray(net["c_bound"], dtype=float).reshape(K))
            n.c_env = f64(np.asarray(net["c_env"], dtype=float).reshape(K, self.E))
            n.D_env = f64(np.asarray(net["D_env"], dtype=float).reshape(K, self.E))
        erx = list(net.get("env_rx_index") or [])
        if erx:
            if not self.is_ecm or env_on is None or not np.any(env_on):
                raise BetseB200Error("extracellular reactions need extracellular spaces and substances that live there")
            ei = capi.as_i32(np.asarray(erx))
            keep.append(ei)
            n.env_rx_prog, n.n_env_rx = ei.ctypes.data_as(C.POINTER(C.c_int32)), len(erx)
            n.stoich_env = f64(np.asarray(net["stoich_env"], dtype=float).reshape(K, len(erx)))
        mods = list(net.get("modulators") or [])
        if mods:
            marr = (capi.Modulator * len(mods))()
            for j, md in enumerate(mods):
                target, prog, mx = md[:3]
                marr[j].target, marr[j].prog, marr[j].max_val = int(target), int(prog), float(mx)
                marr[j].ion = int(md[3]) if len(md) > 3 else -1
            if any(int(md[0]) == 2 for md in mods):
                # tight-junction modulators (networks.py:3301-3317): the squares of the barrier and sim.D_env itself (the
                # device otherwise holds only D_env * TJ_modulator)
                if not self.is_ecm or getattr(self, "_denv", None) is None:
                    raise BetseB200Error("tight-junction modulators need extracellular spaces and sim.D_env")
                tj = capi.as_i32(np.asarray(net["tj_targets"]).reshape(-1))
                keep.append(tj)
                n.tj_targets, n.n_tj = tj.ctypes.data_as(C.POINTER(C.c_int32)), int(tj.size)
                n.D_env_raw = f64(np.asarray(self._denv, dtype=float).reshape(self.I, self.E))
                n.TJ_modulator = f64(np.array(np.broadcast_to(np.asarray(getattr(self, "_tj", 1.0), dtype=float).reshape(-1, self.E)
                                                              if np.ndim(getattr(self, "_tj", 1.0)) else np.asarray(getattr(self, "_tj", 1.0), dtype=float),
                                                              (self.I, self.E)), dtype=float, copy=True))
                self._tj_targets = np.asarray(tj, dtype=np.int64)
            keep.append(marr)
            n.n_modulators, n.modulators = len(mods), marr
        gates = list(net.get("ligand_gates") or [])
        if gates:
            garr = (capi.LigandGate * len(gates))()
            for j, g in enumerate(gates):
                garr[j].species, garr[j].ion, garr[j].extracell = int(g["species"]), int(g["ion"]), int(bool(g["extracell"]))
                garr[j].K, garr[j].n, garr[j].max_val, garr[j].mod = float(g["K"]), float(g["n"]), float(g["max"]), float(g["mod"])
            keep.append(garr)
            n.n_ligand_gates, n.ligand_gates = len(gates), garr
        intra = np.asarray(net.get("intra_on", np.zeros(K)), dtype=np.uint8).reshape(K)
        if intra.any():
            if "mem_vol" not in self.mesh or "R_rads" not in self.mesh:
                raise BetseB200Error("'update intracellular' needs cells.mem_vol and cells.R_rads in the mesh")
            io = np.ascontiguousarray(intra)
            keep.append(io)
            n.intra_on = io.ctypes.data_as(C.POINTER(C.c_uint8))
            n.Do = f64(np.asarray(net["Do"], dtype=float).reshape(K))
            n.c_mems = f64(np.asarray(net["c_mems"], dtype=float).reshape(K, self.M))
            n.R_rads = f64(np.asarray(self.mesh["R_rads"], dtype=float))
            if "mu_mem" in net:
                n.mu_mem = f64(np.asarray(net["mu_mem"], dtype=float).reshape(K))
            if getattr(self, "_Emc0", None) is not None:
                n.Emc = f64(self._Emc0)
            if not trs_needs_vol(net):
                n.mem_sa_over_vol = f64(np.asarray(self.mesh["mem_sa"], dtype=float) / np.asarray(self.mesh["mem_vol"], dtype=float))
        if "map_cell2ecm" in self.mesh and self.is_ecm:
            mc = np.zeros(self.C, dtype=np.int32)
            mc[:self.Co] = np.asarray(self.mesh["map_cell2ecm"]).astype(np.int32)[:self.Co]
            keep.append(mc)
            n.map_cell2ecm = capi.ptr_i32(mc)
        trs = list(net.get("transporters") or [])
        if trs:
            tarr = (capi.Transporter * len(trs))()
            for j, t in enumerate(trs):
                if len(t["terms"]) > capi.TR_MAX_TERMS:
                    raise BetseB200Error("a transporter with more than %d reactants + products" % capi.TR_MAX_TERMS)
                tarr[j].prog, tarr[j].n_terms, tarr[j].net_z = int(t["prog"]), len(t["terms"]), float(t["net_z"])
                for q, (kind, index, sign, coeff) in enumerate(t["terms"]):
                    tt = tarr[j].terms[q]
                    tt.kind, tt.index, tt.sign, tt.coeff = int(kind), int(index), int(sign), float(coeff)
                for member, n_ in (("cell_mask", self.C), ("env_mask", self.E), ("mem_mask", self.M)):
                    mk = t.get(member)
                    if mk is not None:
                        mk = np.ascontiguousarray(np.asarray(mk, dtype=np.uint8).reshape(-1))
                        if mk.size < n_:
                            mk = np.concatenate((mk, np.zeros(n_ - mk.size, dtype=np.uint8)))
                        keep.append(mk)
                        setattr(tarr[j], member, mk.ctypes.data_as(C.POINTER(C.c_uint8)))
            keep.append(tarr)
            n.n_transporters, n.transporters = len(trs), tarr
            if "mem_vol" not in self.mesh:
                raise BetseB200Error("transporters need cells.mem_vol in the mesh")
            n.mem_sa_over_vol = f64(np.asarray(self.mesh["mem_sa"], dtype=float) / np.asarray(self.mesh["mem_vol"], dtype=float))
        pumps = list(net.get("pumps") or [])
        if pumps:
            parr = (capi.SubstancePump * len(pumps))()
            for j, q in enumerate(pumps):
                parr[j].species, parr[j].into_cell, parr[j].uses_ATP = int(q["species"]), int(bool(q["into_cell"])), int(bool(q["uses_ATP"]))
                parr[j].max_val, parr[j].Km = float(q["max"]), float(q["Km"])
            keep.append(parr)
            n.n_pumps, n.pumps = len(pumps), parr
        if net.get("scale_factor") is not None:
            n.scale_factor = f64(np.asarray(net["scale_factor"], dtype=float).reshape(K))
        n.affect_charge = int(bool(self.p.get("substances_affect_charge", 0)) if net.get("affect_charge") is None
                              else bool(net["affect_charge"]))
        self._check(self.lib.betse_set_network(self.ctx, int(handler), C.byref(n)), "betse_set_network")
        self.networks = getattr(self, "networks", {})
        self.networks[int(handler)] = {"species": list(net["species"]), "n_rates": n.n_rates, "env_on": env_on.astype(bool),
                                       "intra_on": intra.astype(bool)}

    def network_state(self, handler=0, rates=False):
        """Substance concentrations [K][C] (and the last rates [n_rates][C]) of a handler."""
        info = self.networks[int(handler)]
        c = np.empty((len(info["species"]), self.C))
        r = np.empty((info["n_rates"], self.C)) if rates else None
        self._check(self.lib.betse_network_state(self.ctx, int(handler), capi.ptr_f64(c),
                                                 capi.ptr_f64(r) if rates else None), "betse_network_state")
        self.d2h_bytes += c.nbytes + (r.nbytes if rates else 0)
        return (c[:, :self.Co], r[:, :self.Co]) if rates else c[:, :self.Co]

    def set_network_events(self, handler, c_bound=None, clamp=None):
        """Scheduled values of the substances' own events for the step about to run (network.event_values)."""
        keep = [None if a is None else capi.as_f64(np.asarray(a, dtype=float)) for a in (c_bound, clamp)]
        self._check(self.lib.betse_network_set_events(self.ctx, int(handler), *(None if a is None else capi.ptr_f64(a) for a in keep)),
                    "betse_network_set_events")

    def set_noise_flux(self, flux, ion="P"):
        """Dynamic noise: the host's draw of sim.protein_noise_flux for the next timestep (sim.py:1322-1339)."""
        if ion not in self.ions:
            raise BetseB200Error("dynamic noise needs the %s ion" % ion)
        a = capi.as_f64(np.asarray(flux, dtype=float).reshape(self.M))
        self.h2d_bytes += a.nbytes
        self._check(self.lib.betse_set_noise_flux(self.ctx, self.ions.index(ion), capi.ptr_f64(a)), "betse_set_noise_flux")

    def network_mem_state(self, handler=0):
        """Membrane values [K][M] of a handler's substances (Molecule.cc_at_mem)."""
        info = self.networks[int(handler)]
        c = np.empty((len(info["species"]), self.M))
        self._check(self.lib.betse_network_mem_state(self.ctx, int(handler), capi.ptr_f64(c)), "betse_network_mem_state")
        self.d2h_bytes += c.nbytes
        return c

    def network_env_state(self, handler=0):
        """Env concentrations [K][E] of a handler's substances (zeros for substances that live in the cells only)."""
        info = self.networks[int(handler)]
        c = np.empty((len(info["species"]), self.E))
        self._check(self.lib.betse_network_env_state(self.ctx, int(handler), capi.ptr_f64(c)), "betse_network_env_state")
        self.d2h_bytes += c.nbytes
        return c

    # ------------------------------------------------------------------ domain decomposition
    def window(self):
        """This rank's exchange window (include/betse_b200.h: betse_window_info)."""
        w = capi.WindowInfo()
        self._check(self.lib.betse_window(self.ctx, C.byref(w)), "betse_window")
        return w

    def set_row_ranges(self, yi, ya, yf):
        self._check(self.lib.betse_set_row_ranges(self.ctx, int(yi[0]), int(yi[1]), int(ya[0]), int(ya[1]),
                                                  int(yf[0]), int(yf[1])), "betse_set_row_ranges")

    def attach_neighbor(self, side, info, plan, same_process):
        """``plan``: dict of the send lists / row blocks towards that neighbour (partition.py)."""
        nb = capi.Neighbor()
        nb.side, nb.same_process, nb.info = int(side), int(bool(same_process)), info
        sc, sf = capi.as_i32(plan["send_cells"]), capi.as_i32(plan["send_flux"])
        nb.n_send_cells, nb.send_cells, nb.recv_cell0 = len(sc), capi.ptr_i32(sc), int(plan["recv_cell0"])
        nb.n_send_flux, nb.send_flux, nb.recv_slot0 = len(sf), capi.ptr_i32(sf), int(plan["recv_slot0"])
        nb.cc_rows, nb.cc_src_row0, nb.cc_dst_row0 = (int(x) for x in plan["cc_rows"])
        nb.v_rows, nb.v_src_row0, nb.v_dst_row0 = (int(x) for x in plan["v_rows"])
        self._check(self.lib.betse_attach_neighbor(self.ctx, C.byref(nb)), "betse_attach_neighbor")

    def exchange(self, which, buf_next, mode):
        self._check(self.lib.betse_exchange(self.ctx, int(which), int(bool(buf_next)), int(mode)), "betse_exchange")

    def step_phase(self, phase, diag=False):
        self._check(self.lib.betse_step_phase(self.ctx, int(phase), capi.STEP_DIAG if diag else 0), "betse_step_phase")

    def update_V_phase(self, phase):
        self._check(self.lib.betse_update_v_phase(self.ctx, int(phase)), "betse_update_v_phase")

    def sync(self):
        st = C.c_uint32(0)
        self._check(self.lib.betse_sync(self.ctx, C.byref(st)), "betse_sync")
        return int(st.value)

    def close(self):
        pf = self.__dict__.pop("_prefetch", None)
        if pf is not None:
            pf.release()
        self.__dict__.pop("_pinned", {}).clear()        # staging not lent to anyone is freed here (see _pinned_array)
        if getattr(self, "ctx", None):
            self.lib.betse_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class EnsembleEngine:
    """B independent small tissues on ONE GPU, advanced together (last bullet: tissues of 228-10 k cells are
    launch-latency bound, the reference runs one process per parameter set, simrunner.py:93-296).  Every member is a full
    TissueEngine of its own — its own state, parameters, schedule, channels — so a member computes exactly what it would
    compute alone (bit-identical, tests/test_gpu_ensemble.py); `step` hands all members to betse_ensemble_step, which
    runs n x k timesteps as one CUDA graph with one stream per member.

    ``members``: list of (mesh, params, state) triples, or ``replicas=B`` copies of one triple (a parameter ensemble
    starts from copies and then calls ``members[j].set_field`` / ``upload`` per member)."""

    STEPS_PER_LAUNCH = 10

    def __init__(self, mesh=None, params=None, state=None, device=0, replicas=1, members=None):
        if members is None:
            members = [(mesh, params, state)] * int(replicas)
        self.members = [TissueEngine(m, p, s, device=device) for (m, p, s) in members]
        self.lib = self.members[0].lib
        self.B = len(self.members)
        self._ctxs = (C.c_void_p * self.B)(*[e.ctx for e in self.members])
        self._warm = False
        self.device_ms = 0.0
        m0 = self.members[0]
        self.Co, self.M, self.E, self.I = m0.Co, m0.M, m0.E, m0.I

    @property
    def h2d_bytes(self):
        return sum(e.h2d_bytes for e in self.members)

    @property
    def d2h_bytes(self):
        return sum(e.d2h_bytes for e in self.members)

    def update_V(self):
        for e in self.members:
            e.update_V()

    def set_channels(self, specs, **kw):
        for e in self.members:
            e.set_channels([dict(s) for s in specs], **kw)

    def _run(self, k, launches):
        st = (C.c_uint32 * self.B)()
        ms = C.c_float(0)
        rc = self.lib.betse_ensemble_step(self._ctxs, self.B, int(k), int(launches), st, C.byref(ms))
        self.members[0]._check(rc, "betse_ensemble_step")
        for e in self.members:
            e.steps_done += k * launches
        self.device_ms += float(ms.value)
        return [int(x) for x in st], float(ms.value)

    def step(self, n=1):
        """n timesteps of every member; returns the OR of the members' status words (``self.status`` holds them all)."""
        status = [0] * self.B
        n = int(n)
        if n > 0 and not self._warm:
            status = [e.step(1) for e in self.members]          # loads every kernel outside a capture
            self._warm = True
            n -= 1
        k = self.STEPS_PER_LAUNCH
        for chunk, count in ((k, n // k), (n % k, 1 if n % k else 0)):
            if count:
                st, _ = self._run(chunk, count)
                status = [a | b for a, b in zip(status, st)]
        self.status = status
        out = 0
        for s in status:
            out |= s
        return out

    def profile(self, n):
        """n timesteps of every member, device-timed; -> (total ms, {})."""
        if not self._warm:
            self.step(1)
        k = min(self.STEPS_PER_LAUNCH, n)
        self._run(k, 2)                                        # capture + instantiate (both parities) outside the timed launches
        _, ms = self._run(k, n // k)
        if n % k:
            self._run(n % k, 1)
            ms += self._run(n % k, 1)[1]
            ms -= 0.0
        return ms, {}

    def download(self, fields, pinned=False):
        """Fields of every member stacked on a leading member axis."""
        got = [e.download(fields, pinned=pinned) for e in self.members]
        return {f: np.stack([g[f] for g in got]) for f in got[0]}

    def close(self):
        for e in self.members:
            e.close()
