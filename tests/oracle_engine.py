"""TEST INFRASTRUCTURE: a stand-in for betse_b200.engine.TissueEngine backed by the CPU oracle.

The drop-in loop (betse_b200/simloop.py) is host logic around the CUDA engine: event scheduling, the cutting event,
sampling, copy-back, write2storage.  The build container has the real reference but no GPU, the GPU box has a GPU but
no reference — so the host logic is held to the reference's own full run HERE, with the device replaced by the
oracle (tests/test_simloop_reference.py), and the device is held to the oracle / the recorded reference on the GPU
box (tests/test_gpu_*.py).  Never imported by the product."""
import numpy as np

from oracle.betse_oracle import OracleFastSim, OracleNetwork, OracleSim, SimUnstable


class OracleEngine:
    def __init__(self, mesh, params, state, device=0, partition=None):
        self._args = (mesh, params, state)
        self.ions = [str(x) for x in params["ions"]]
        self.Co = self.C = len(mesh["cell_vol"])
        self.M = len(mesh["mem_sa"])
        self.is_ecm = bool(params["is_ecm"])
        self.ny, self.nx = (int(x) for x in mesh["grid_shape"])
        self.mem_to_cells = np.asarray(mesh["mem_to_cells"])
        self.h2d_bytes = self.d2h_bytes = 0
        self.noecm_field = (not self.is_ecm) and "D_env_weight" in state
        self.networks = {}
        self._descs = {}
        self._specs = []
        self._phase_init = False
        self._o = None
        self.sets = []

    # ---- what simloop.engine_from_sim calls
    def set_network(self, comp, handler=0):
        self._descs[int(handler)] = comp["_desc"]          # attached by the test's compile_network wrapper
        self.networks[int(handler)] = {"species": list(comp["species"]),
                                       "env_on": np.asarray(comp.get("env_on", np.zeros(len(comp["species"]))), dtype=bool),
                                       "intra_on": np.asarray(comp.get("intra_on", np.zeros(len(comp["species"]))), dtype=bool)}

    def set_channels(self, specs, phase_init=False, affect_charge=None):
        self._specs = [{k: v for k, v in c.items() if k != "_obj"} for c in specs]
        self._phase_init = bool(phase_init)

    def _sim(self):
        if self._o is None:
            mesh, params, state = self._args
            state = dict(state)
            # sim.py:781-786 (the engine derives these from the mesh too)
            nm = np.asarray(mesh["num_mems"], dtype=float)[np.asarray(mesh["mem_to_cells"])]
            nfrac = float(params["smooth_cells"])
            state.setdefault("smooth_weight_mem", (nfrac * nm - 1) / (nfrac * nm))
            state.setdefault("smooth_weight_o", 1 / (nfrac * nm))
            # no-ECM field diagnostics (ion_current.py:116-171) are not part of the engine's contract: any weight will do
            state.setdefault("D_env_weight", np.ones(self.ny * self.nx))
            self._o = OracleSim(mesh, params, state, channels=self._specs, phase_init=self._phase_init,
                                networks=[self._descs[h] for h in sorted(self._descs)], net_handlers=sorted(self._descs))
            self._active = [c for c in self._o.channels if not (self._phase_init and not c["init_active"])]
        return self._o

    def set_field(self, name, value):
        self.sets.append(name)
        if getattr(self, "_f", None) is not None and name == "gj_block":
            self._f.gj_block = np.array(value, dtype=float)
            return
        o = self._sim()
        if name == "bound_V":
            o.bound_V = dict(zip("TBLR", value))
        else:
            setattr(o, name, np.array(value, dtype=float))

    def set_bound_V(self, v):
        self.set_field("bound_V", v)

    def set_bath(self, values):
        self.sets.append("bath")
        o = self._sim()
        for i, v in values.items():
            o.cc_env[int(i)][:] = float(v)

    def step(self, n=1, diag=False):
        o = self._sim()
        o.diagnostics = True
        for _ in range(n):
            try:
                o.step()
            except SimUnstable:
                return 1
        return 0

    def download(self, fields, pinned=False):
        o = self._sim()
        out = {}
        for f in fields:
            if not self.is_ecm and (f == "rho_env" or (f in ("E_env_x", "E_env_y", "v_env") and not self.noecm_field)):
                continue
            out[f] = np.array(getattr(o, f), dtype=float, copy=True)
            if out[f].ndim == 2 and f != "cc_cells" and out[f].shape == (self.ny, self.nx):
                out[f] = out[f].ravel()          # the engine hands env-grid fields out flat
        return out

    def channel_state(self, k):
        c = self._active[k]
        m, h = np.ones(self.M), np.ones(self.M)
        m[c["targets"]], h[c["targets"]] = c["m"], c["h"]
        return {"m": m, "h": h, "P": c["P"], "flux": c.get("flux", np.zeros(self.M)),
                "DChan": c.get("DChan", np.zeros(self.M))}

    def network_state(self, handler=0, rates=False):
        net = self._sim().networks[sorted(self._descs).index(int(handler))]
        c = np.stack([net.c[n] for n in net.species])
        if rates:
            r = net.rates if net.rates is not None else np.zeros((len(net.species), self.C))
            return c, np.asarray(r)
        return c

    def set_noise_flux(self, flux, ion="P"):
        self.sets.append("noise")
        self._sim().noise_flux = np.array(flux, dtype=float)

    def set_network_events(self, handler, c_bound=None, clamp=None):
        self.sets.append("net_events")
        net = self._sim().networks[sorted(self._descs).index(int(handler))]
        net.forced_events = (np.array(c_bound, dtype=float), np.array(clamp, dtype=float))

    def tj_modulator(self):
        if not any(len(np.asarray(d.get("tj_targets", []))) for d in self._descs.values()):
            return None
        return np.array(self._sim().TJ_modulator, dtype=float, copy=True)

    def network_mem_state(self, handler=0):
        net = self._sim().networks[sorted(self._descs).index(int(handler))]
        return np.stack([net.cmem[n] for n in net.species])

    def network_env_state(self, handler=0):
        net = self._sim().networks[sorted(self._descs).index(int(handler))]
        return np.stack([net.c_env.get(n, np.zeros(self.ny * self.nx)) for n in net.species])

    # ---- fast (equivalent-circuit) solver
    FAST_FIELDS = {"vm_ave": "C", "gjopen": "M", "vgj": "M", "Jn": "M", "Emx": "M", "Emy": "M",
                   "J_cell_x": "C", "J_cell_y": "C", "E_cell_x": "C", "E_cell_y": "C"}

    def fast_set_channels(self, cbar, rev_E, geo_conv=1.0):
        self._fast_consts = {"cbar": np.asarray(cbar, dtype=float), "rev_E": np.asarray(rev_E, dtype=float),
                             "geo_conv": float(geo_conv), "zs": np.asarray(self._args[2]["zs"], dtype=float)}

    def fast_setup(self, state):
        mesh, params, st = self._args
        specs = [dict(c, targets=np.arange(self.M) if c.get("targets") is None else c["targets"]) for c in self._specs]
        self._f = OracleFastSim(mesh, params, dict(st, **{k: v for k, v in state.items() if v is not None}),
                                channels=specs, phase_init=self._phase_init, fast_consts=getattr(self, "_fast_consts", None))
        self._active = [c for c in self._f.channels if not (self._phase_init and not c["init_active"])]

    def fast_step(self, n=1, diag=False):
        for _ in range(n):
            try:
                self._f.step()
            except FloatingPointError:
                return 1
        return 0

    def fast_download(self, fields=("vm_ave", "gjopen")):
        out = {f: np.array(getattr(self._f, f), dtype=float, copy=True) for f in fields}
        if "vm_ave" in out:
            out["vm"] = out["vm_ave"][np.asarray(self.mem_to_cells).astype(np.int64)]
        return out

    def close(self):
        pass
