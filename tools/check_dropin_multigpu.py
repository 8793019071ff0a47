#!/usr/bin/env python3
"""The REAL drop-in over N GPUs (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 \
        tools/check_dropin_multigpu.py

Every rank runs the unmodified reference's SimRunner.seed/init/sim (baseline/_ref or /root/reference through
oracle/refshim.py) with `simloop.install()`: inside an initialised process group run_sim_core_loop steps the tissue as N
strips (betse_b200/simloop.py:_run_strips).  Rank 0 then runs the same configuration with the reference's own NumPy loop
and compares the stored time series.  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def run(tmp, use_dropin, mods):
    from oracle import refrun, refshim
    refshim.bypass_science_init()
    from betse.science.parameters import Parameters
    from betse.science.simrunner import SimRunner
    from betse.science.phase import phasecallbacks
    from betse_b200 import simloop
    if use_dropin:
        simloop.install()
    try:
        fn = refrun.write_config(tmp, mods)
        np.random.seed(12345)
        p = Parameters.make(fn)
        p.anim.is_while_sim = p.anim.is_after_sim = p.plot.is_after_sim = False
        runner = SimRunner(p=p, callbacks=phasecallbacks.SimCallbacksNoop())
        runner.seed()
        runner.init()
        phase = runner.sim()
    finally:
        if use_dropin:
            simloop.uninstall()
    return phase.sim


def main():
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from tests.golden.make_golden import NO_NET, _m
    # the shipped world (228 cells, 24 x 25 env grid) with the mammal ion profile, ion path only
    mods = _m(NO_NET, {"general options": {"ion profile": "mammal"}})
    with tempfile.TemporaryDirectory() as d:
        new = run(os.path.join(d, "new"), True, mods)
        ok, worst = True, {}
        if rank == 0:
            os.environ["BETSE_STRIPS"] = "0"
            ref = run(os.path.join(d, "ref"), False, mods)
            for name in ("vm_time", "cc_time", "cc_env_time", "gjopen_time", "vm_ave_time", "rho_cells_time", "I_mem_time", "venv_time"):
                got, want = getattr(new, name), getattr(ref, name)
                same_len = len(got) == len(want) and len(want) >= 30
                err = max(float(np.max(np.abs(np.asarray(a, dtype=float) - np.asarray(r, dtype=float)))
                                / max(float(np.max(np.abs(np.asarray(r, dtype=float)))), 1e-300)) for a, r in zip(got, want)) if same_len else float("inf")
                worst[name] = err
                # I_mem is a difference of cancelling membrane currents (a diagnostic): judged an order looser
                ok &= same_len and err <= (1e-7 if name == "I_mem_time" else 1e-8)
            print(json.dumps({"check": "drop-in over %d strips == reference loop" % dist.get_world_size(), "ok": bool(ok),
                              "world": dist.get_world_size(), "max_rel_err": worst}))
    flag = [ok]
    dist.broadcast_object_list(flag, src=0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag[0] else 1)


if __name__ == "__main__":
    main()
