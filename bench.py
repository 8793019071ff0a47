#!/usr/bin/env python3
"""bench.py — timesteps/s and membrane-updates/s of the tissue-update loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1..c5] [--cells C] [--impl ours|reference]

One "step" = one timestep of Simulator._run_sim_core_loop (betse/science/sim.py:1169-1365) over the whole tissue.
Workload at N=1 (default, --config c5): BASELINE.json configs[4] on one GPU — a 1 M-cell (6 M-membrane) mammal-profile
tissue with extracellular spaces on a ~1000x1000 grid.  The other BASELINE configs are bench lines of their own:

    c1  configs[0]  the reference's own 228-cell cluster (mesh built by Cells.make_world, recorded in
                    tests/golden/mammal_noecm.npz), Na/K/Cl/Ca/P/M, no extracellular spaces
    c2  configs[1]  10 k cells + extracellular grid
    c3  configs[2]  100 k cells + Nav1p3 / Kv1p5 / KLeak / Cav1p2 channels + pumps
    c4  configs[3]  50 k cells + the shipped gene regulatory network (extra_configs/grn_basic.yaml)
    --ensemble B    (c1/c2) B independent replicas of the tissue advanced by ONE engine (small tissues run
                    as parameter ensembles, one per GPU under torchrun)

Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "membrane_updates_per_sec"
UNIT = "membrane-updates/s"
KNAMES = ["k_ion", "k_mem", "k_envacc", "k_field", "k_envmix", "k_ion_smooth", "k_diag", "k_xchg"]
BASELINE_CONFIG = {"c1": "configs[0]", "c2": "configs[1]", "c3": "configs[2]", "c4": "configs[3]", "c5": "configs[4] on one GPU"}


def algorithmic_bytes(I, C, M, E, ecm=True, n_chan=0, n_subst=0):
    """every persistent array any implementation must read/write once per step.
    Returns (whole step, membrane/cell kernel share, env kernels share).  Per voltage-gated channel on all
    membranes: 32 M (m, h read + written) + 8 M (P); per network substance: 32 C."""
    mem = I * (32 * C + 8 * M) + 48 * M + 48 * C + n_chan * 40 * M + n_subst * 32 * C
    env = (I * 24 * E + 72 * E) if ecm else 0
    return mem + env, mem, env


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                     timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(s[0]) for s in self.samples)
        reasons = []
        for j, name in enumerate(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]):
            if any(s[3 + j].lower().startswith("active") for s in self.samples):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]),
                "power_w_max": max(float(s[2]) for s in self.samples), "reasons": reasons,
                "samples": len(self.samples)}


def measured_peak_hbm():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        try:
            return float(json.load(open(fn))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class NS:
    """Duck-typed stand-in for the reference's Simulator / Cells / Parameters objects."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def namespaces(mesh, p, state, kind="INIT"):
    cells = NS(**{k: v for k, v in mesh.items()})
    cells.mem_nx, cells.mem_ny = mesh["mem_nx"], mesh["mem_ny"]
    cells.grid_shape = tuple(int(x) for x in mesh["grid_shape"])
    cells.X = None
    pp = NS(**{k: v for k, v in p.items()})
    sim = NS(**{k: (np.array(v, copy=True) if isinstance(v, np.ndarray) else v) for k, v in state.items()})
    sim.bound_V = {"T": 0, "B": 0, "L": 0, "R": 0}
    sim.sampled = 0
    sim.write2storage = lambda t, cells, p: setattr(sim, "sampled", sim.sampled + 1)
    phase = NS(p=pp, cells=cells, sim=sim, kind=NS(name=kind), callbacks=NS(progressed_next=lambda: None))
    if kind == "SIM":
        # a SIM phase calls TissueHandler.fire_events every step (sim.py:1187-1188): here one that schedules nothing, under
        # a configuration without scheduled interventions (every entry of p.global_options / p.scheduled_options off)
        phase.dyna = NS(fire_events=lambda phase, t: None, event_cut=None)
        pp.global_options = {k: 0 for k in ("K_env", "Cl_env", "Na_env", "T_change", "gj_block", "NaKATP_block")}
        pp.scheduled_options = {k: 0 for k in ("Na_mem", "K_mem", "Cl_mem", "Ca_mem", "pressure", "ecmJ", "cuts")}
    return sim, phase


# ---------------------------------------------------------------------------- workloads
def build_workload(cfg, cells=None, profile="mammal", ecm=True, ragged=0.0):
    """-> dict(mesh, p, state, kind of extras).  Channel gate states / network tables are attached by `attach`."""
    from betse_b200 import synth
    w = {"config": cfg, "channels": False, "network": None}
    if cfg == "c1":
        from tests import util
        cap = util.load_golden("mammal_noecm")
        w["mesh"], w["p"], w["state"] = util.group(cap, "cells."), util.group(cap, "sim.p."), util.group(cap, "sim.s0.")
        w["what"] = "the reference's own %d-cell cluster (Cells.make_world, seed 12345; tests/golden/mammal_noecm.npz)" % len(w["mesh"]["cell_vol"])
        return w
    n = {"c2": 10_000, "c3": 100_000, "c4": 50_000, "c5": 1_000_000}[cfg] if cells is None else int(cells)
    mesh, p, state = synth.make_tissue(n, profile=profile, ecm=ecm, dt=1.0e-4, ragged=ragged)
    w["mesh"], w["p"], w["state"] = mesh, p, state
    w["ragged"] = ragged
    if cfg == "c3":
        p["substances_affect_charge"] = 1
        w["channels"] = True
    if cfg == "c4":
        from betse_b200 import network as netlib
        from tests import util
        cap = util.load_golden("mammal_ecm_grn")
        desc = netlib.unflatten(cap, "sim.s0.net1.")
        C, M, E = len(mesh["cell_vol"]), len(mesh["mem_sa"]), int(np.prod(mesh["grid_shape"]))
        w["network"] = (synth.retarget_network(desc, C, M, E, np.random.default_rng(7)), 1)
        p["substances_affect_charge"] = int(cap["sim.p.substances_affect_charge"])
    return w


def attach(eng, w, vm=None):
    """Channels / network of the workload onto a TissueEngine (after update_V) — or, with ``vm`` given, just the specs."""
    from betse_b200 import network as netlib
    from betse_b200 import synth
    out = {"specs": [], "n_chan": 0, "n_subst": 0}
    if w["network"] is not None:
        desc, h = w["network"]
        out["n_subst"] = len(desc["species"])
        if eng is not None:
            eng.set_network(netlib.compile_network(desc, eng.Co, eng.M), handler=h)
            eng.set_channels([])
    if w["channels"]:
        v = vm if vm is not None else eng.download(["vm"])["vm"]
        out["specs"] = synth.baseline_channels(v)
        out["n_chan"] = len(out["specs"])
        if eng is not None:
            eng.set_channels([dict(s) for s in out["specs"]])
    return out


def cpu_reference_run(w, budget_s, max_steps):
    """Times the oracle (the CPU restatement of the reference loop) on the same workload."""
    from oracle.betse_oracle import OracleSim
    mesh, p, state = w["mesh"], w["p"], w["state"]
    kw = {}
    if w["channels"]:
        o0 = OracleSim(mesh, p, state)
        o0.diagnostics = False
        o0.update_V()
        kw["channels"] = attach(None, w, vm=o0.vm)["specs"]
    if w["network"] is not None:
        kw["networks"], kw["net_handlers"] = [w["network"][0]], [w["network"][1]]
    o = OracleSim(mesh, p, state, **kw)
    o.diagnostics = False
    o.update_V()
    t0 = time.perf_counter()
    o.step()
    first = time.perf_counter() - t0
    n = int(max(1, min(max_steps, budget_s / max(first, 1e-9))))
    t0 = time.perf_counter()
    for _ in range(n):
        o.step()
    dt = (time.perf_counter() - t0) / n
    try:
        from threadpoolctl import threadpool_info
        thr = max([i.get("num_threads", 1) for i in threadpool_info()] + [1])
    except Exception:
        thr = 1
    return dt, n, thr


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--config", default="c5", choices=sorted(BASELINE_CONFIG))
    ap.add_argument("--cells", type=int, default=None, help="override the synthetic tissue size of c2..c5")
    ap.add_argument("--ensemble", type=int, default=0, help="replicas of a small tissue advanced by one engine (c1/c2)")
    ap.add_argument("--ragged", type=float, default=0.0, help="fraction of gap-junction membrane pairs removed from the synthetic sheet (cells with 3-6 membranes)")
    ap.add_argument("--profile", default="mammal")
    ap.add_argument("--no-ecm", action="store_true")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for cpu_baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-kind", default="INIT", choices=["INIT", "SIM"],
                    help="phase kind of the end-to-end run: SIM calls a (no-op) fire_events before every step, one step per call")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    cfg = args.config
    w = build_workload(cfg, args.cells, args.profile, not args.no_ecm, args.ragged)
    mesh, p, state = w["mesh"], w["p"], w["state"]
    ecm = bool(p["is_ecm"])
    I = len(p["ions"])
    C, M = len(mesh["cell_vol"]), len(mesh["mem_sa"])
    gny, gnx = (int(x) for x in mesh["grid_shape"])
    E = gny * gnx
    B = max(1, args.ensemble)
    n_chan = len(__import__("betse_b200.synth", fromlist=["x"]).BASELINE_CHANNELS) if w["channels"] else 0
    n_subst = len(w["network"][0]["species"]) if w["network"] is not None else 0
    workload = w.get("what") or "%s-cell synthetic tissue (%d cells, %d membranes), %s ion profile (I=%d), %s%s%s" % (
        "1M" if abs(C - 1e6) < 2e4 else str(C), C, M, args.profile, I,
        "extracellular grid %dx%d" % (gny, gnx) if ecm else "no extracellular spaces",
        ", %d voltage-gated channels on every membrane (Nav1p3, Kv1p5, KLeak, Cav1p2)" % n_chan if n_chan else "",
        ", gene regulatory network of %d substances (grn_basic.yaml)" % n_subst if n_subst else "")
    if w.get("ragged"):
        workload += ", ragged: %.0f %% of the gap-junction membrane pairs removed (3-6 membranes per cell)" % (100 * w["ragged"])
    if B > 1:
        workload += " x %d independent replicas in one engine" % B
    b_step, b_mem, b_env = algorithmic_bytes(I, C * B, M * B, E * B, ecm, n_chan, n_subst)
    config = {"decomposition": ("%d strips of env-grid rows, halo exchange by CUDA-IPC peer stores over NVLink (no NCCL on the data path)" % world)
              if (world > 1 and cfg == "c5") else ("one replica per GPU" if world > 1 else "single domain"),
              "workload": workload, "baseline_config": BASELINE_CONFIG[cfg] if (args.cells is None and not args.ragged) else "custom",
              "cells": C * B, "membranes": M * B, "env_points": E * B if ecm else 0, "ions": I, "dt": float(p["dt"]),
              "l2_policy": ("state per step (~%.2f GB) is larger than the 126 MB L2" % (b_step / 1e9)) if b_step > 2.5e8 else
                           ("state per step is %.1f MB: L2-resident — a write of a 256 MB scratch buffer flushes the L2 between timed batches" % (b_step / 1e6)),
              "sampling": "none inside `value`; every 10 steps inside `e2e`"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        per_step, n, thr = cpu_reference_run(w, budget_s=60.0, max_steps=max(1, args.steps))
        val = M / per_step
        line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
                "steps": n, "steps_requested": args.steps, "warmup": 1, "ms_per_step": per_step * 1e3,
                "timesteps_per_sec": 1.0 / per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": val, "unit": UNIT, "cores": thr, "kind": "port",
                                 "sample": "%d timesteps (after 1 warm-up step; as many of the %d requested as fit 60 s) of the same "
                                           "workload with the NumPy oracle (sparse restatement of the reference loop; the reference's "
                                           "dense operators need 48*C^2 bytes and cannot hold this tissue)" % (n, args.steps)},
                "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm (GPU)
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))      # plumbing only: handles, barriers, max-reduce of times
    from betse_b200.engine import EnsembleEngine, TissueEngine
    from betse_b200 import simloop

    strips = world > 1 and cfg == "c5"
    ds = None
    if strips:
        # one tissue cut into `world` strips (strong scaling): halo exchange over NVLink peer stores
        from betse_b200.strips import DistributedStrips, verify_strips
        # parity first: a small tissue cut into the same number of strips must equal the undivided one bit for bit
        strips_parity = verify_strips(dist, local_rank, cells=60_000, steps=12)
        verdict = [strips_parity["ok"] if rank == 0 else None]
        dist.broadcast_object_list(verdict, src=0)
        if not verdict[0]:
            if rank == 0:
                print(json.dumps({"error": "decomposed tissue differs from the undivided one", "strips_parity": strips_parity}))
            dist.destroy_process_group()
            sys.exit(1)
        ds = DistributedStrips(mesh, p, state, local_rank, dist)
        eng = ds.engine
        ds.update_V()
        ds.step(args.warmup)
    else:
        # N > 1 with a small tissue: one independent replica per GPU, no communication
        eng = EnsembleEngine(mesh, p, state, device=local_rank, replicas=B) if B > 1 else TissueEngine(mesh, p, state, device=local_rank)
        eng.update_V()
        attach(eng, w)
        eng.step(args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    if b_step <= 2.5e8:
        # L2-resident workload: flush the L2 before the timed region (timing rule); the steps themselves re-touch their
        # state every timestep by construction, which is what the metric measures
        scratch = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        scratch.fill_(1)
        torch.cuda.synchronize()
    total_ms, kms = eng.profile(args.steps)
    torch.cuda.synchronize()
    if dist:
        t = torch.tensor([total_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    sampler.stop_flag = True
    status = eng.step(0)
    check = eng.download(["vm", "cc_cells"])
    finite = bool(np.isfinite(check["vm"]).all() and np.isfinite(check["cc_cells"]).all())
    if dist:
        t = torch.tensor([0.0 if finite else 1.0, float(status)], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        finite, status = bool(t[0].item() == 0.0), int(t[1].item())
        kt = torch.tensor([kms.get(k, 0.0) for k in KNAMES], device="cuda")
        dist.all_reduce(kt, op=dist.ReduceOp.MAX)
        kms = {k: float(v) for k, v in zip(KNAMES, kt.tolist()) if v > 0}
    solo_ms = None
    if B > 1:
        # kernel shares and the step time of ONE member stepped alone (after the timed region; the ensemble is done)
        solo_total, kms = eng.members[0].profile(args.steps)
        solo_ms = solo_total / args.steps
    if ds is not None:
        ds.close()
    else:
        eng.close()

    ms_per_step = total_ms / args.steps
    steps_per_s = 1e3 / ms_per_step
    # strips: ONE tissue over N GPUs (strong scaling); replicas: N (x B) independent tissues (weak scaling)
    n_tissues = B * (1 if strips else world)
    value = M * n_tissues * steps_per_s

    # ---- end-to-end through the public API with host buffers
    e2e = None
    if not args.no_e2e and strips:
        # end to end over N strips: host state -> strips (partition on the host, upload), timesteps,
        # download of each rank's owned Vmem / concentrations every 10 steps
        from betse_b200.strips import DistributedStrips
        n_e2e = args.steps
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ds2 = DistributedStrips(mesh, p, state, local_rank, dist)
        ds2.update_V()
        t_build = time.perf_counter() - t0
        done = 0
        d2h = 0
        while done < n_e2e:
            k = min(10, n_e2e - done)
            ds2.step(k)
            done += k
            out = ds2.download_local(["vm", "cc_cells", "cc_env", "gjopen"], pinned=True)
            d2h += sum(a.nbytes for a in out.values())
        torch.cuda.synchronize()
        dist.barrier()
        wall = time.perf_counter() - t0
        t_loop = wall - t_build
        tt = torch.tensor([wall, float(ds2.engine.h2d_bytes), float(d2h)], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ds2.close()
        e2e = {"value": M * n_e2e / float(tt[0].item()), "unit": UNIT,
               "h2d_bytes_per_step": float(tt[1].item()) / n_e2e, "d2h_bytes_per_step": float(tt[2].item()) / n_e2e,
               "timesteps": n_e2e, "sampled_steps": n_e2e // 10,
               "what": "DistributedStrips from host NumPy state on every rank: partition + engine creation + "
                       "strip upload + timesteps + download of the rank's owned vm/cc_cells/cc_env/gjopen every "
                       "10 steps into page-locked staging (bytes are per rank, max over ranks)",
               "seconds": {"partition_engine_upload": round(t_build, 4), "steps_and_samples": round(t_loop, 4)}}
    elif not args.no_e2e and (w["channels"] or w["network"] is not None or B > 1):
        # channels / networks / ensembles: the engine's public API from host NumPy state (the drop-in loop builds the same
        # calls from live reference objects, which the GPU box does not have for a synthetic tissue)
        n_e2e = args.steps
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng2 = EnsembleEngine(mesh, p, state, device=local_rank, replicas=B) if B > 1 else TissueEngine(mesh, p, state, device=local_rank)
        eng2.update_V()
        attach(eng2, w)
        done = 0
        while done < n_e2e:
            k = min(10, n_e2e - done)
            eng2.step(k)
            done += k
            eng2.download(["vm", "cc_cells", "cc_env", "gjopen"], pinned=True)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        h2d, d2h = eng2.h2d_bytes, eng2.d2h_bytes
        eng2.close()
        if dist:
            t = torch.tensor([wall], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": M * n_tissues * n_e2e / wall, "unit": UNIT, "h2d_bytes_per_step": h2d / n_e2e,
               "d2h_bytes_per_step": d2h / n_e2e, "timesteps": n_e2e, "sampled_steps": n_e2e // 10,
               "what": "TissueEngine public API from host NumPy state: engine creation + state upload + channel / network "
                       "set-up + timesteps + download of vm/cc_cells/cc_env/gjopen every 10 steps into page-locked staging"}
    elif not args.no_e2e:
        sim, phase = namespaces(mesh, p, state, kind=args.e2e_kind)
        n_e2e = args.steps
        ts = np.linspace(0, n_e2e * p["dt"], n_e2e)
        sampled = set(ts[10::10].tolist())
        stats = {}
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        simloop.run_sim_core_loop(sim, phase, ts, sampled, None, device=local_rank, stats=stats)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        t_close = time.perf_counter()
        simloop._join_closing()                       # the engine's teardown runs on a helper thread: reported, not hidden
        t_close = time.perf_counter() - t_close
        if dist:
            t = torch.tensor([wall], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        e2e = {"value": M * n_tissues * n_e2e / wall, "unit": UNIT,
               "h2d_bytes_per_step": stats["h2d_bytes"] / n_e2e, "d2h_bytes_per_step": stats["d2h_bytes"] / n_e2e,
               "timesteps": n_e2e, "sampled_steps": sim.sampled, "seconds": stats.get("seconds"),
               "phase_kind": args.e2e_kind,
               "teardown_seconds_not_in_value": round(t_close, 4),
               "value_with_teardown": M * n_tissues * n_e2e / (wall + t_close),
               "what": "run_sim_core_loop() (the Simulator._run_sim_core_loop drop-in) from host NumPy state: "
                       "engine creation + full state upload + timesteps + download of every write2storage "
                       "attribute at each sampled step and at the end; the engine's teardown (device frees, unpinning) runs "
                       "on a helper thread after the loop has returned and is reported separately"}

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_hbm()
    dom = max((k for k in kms if k != "k_xchg"), key=lambda k: kms[k])
    # per-launch algorithmic bytes of one rank's kernel (a rank holds 1/world of a decomposed tissue)
    div = world if strips else (B if B > 1 else 1)      # ensembles: kernel_ms are of ONE member
    share = {"k_mem": b_mem / div, "k_ion": I * 24 * E * B / div, "k_envacc": 0, "k_field": 72 * E * B / div, "k_envmix": 0}
    ach = share.get(dom, b_mem / div) / (kms[dom] * 1e-3) / 1e9
    traffic = None     # measured DRAM bytes per launch of the dominant kernel (one `ncu --set full` capture of this workload)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        wl = tr["workload"]
        if wl["cells"] == C and wl["ions"] == I and bool(wl["ecm"]) == bool(ecm) and world == wl["n_gpus"] and cfg == "c5":
            traffic = tr["bytes_per_launch"].get(dom)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
            "traffic": traffic, "traffic_source": "profiles/ncu_traffic.json (one ncu --set full capture of this workload; not re-measured in this run)" if traffic else None,
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": share.get(dom, b_mem / div),
            "kernel_ms": kms, "step": {"algorithmic_bytes": b_step, "n_gpus": world,
                                       "achieved": b_step / (ms_per_step * 1e-3) / 1e9 * (1 if strips else world),
                                       "frac": b_step / (ms_per_step * 1e-3) / 1e9 / (peak * (world if strips else 1))}}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "timesteps_per_sec": steps_per_s,
            "higher_is_better": True, "scaling": "strong" if strips else ("weak" if world > 1 else "strong"), "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config, "clocks": sampler.summary(),
            "gpu_launches": int((len(kms) + (1 if "k_xchg" in kms else 0)) * args.steps * world * B),
            "status_word": status, "finite": finite,
            **({"strips_parity": strips_parity} if strips else {}),
            **({"ensemble": {"members": B, "solo_ms_per_step": solo_ms, "speedup_vs_solo_per_gpu": B * solo_ms / ms_per_step,
                             "what": "B independent tissues, one CUDA graph of B streams x 10 timesteps per launch (betse_ensemble_step); "
                                     "kernel_ms are of one member stepped alone"}} if B > 1 else {}),
            "roofline": roof}
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        per_step, n, thr = cpu_reference_run(w, budget_s=args.cpu_budget, max_steps=20)
        line["cpu_baseline"] = {"value": M / per_step, "unit": UNIT, "cores": thr, "kind": "port",
                                "ms_per_step": per_step * 1e3,
                                "sample": "%d timesteps of ONE %d-cell tissue of the same workload with the NumPy oracle" % (n, C)}
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
