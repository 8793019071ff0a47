#!/usr/bin/env python3
"""Multi-process / multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_multigpu.py [--cells 60000] [--steps 12]

Every rank steps its strip of ONE tissue with halo exchange over CUDA-IPC peer stores
(betse_b200/strips.py: DistributedStrips); rank 0 also runs the undivided tissue on its GPU and
the gathered strips must equal it BIT-EXACTLY.  Prints one JSON line on rank 0; exit code 1 on
mismatch."""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=60_000)
    ap.add_argument("--steps", type=int, default=12)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from betse_b200 import synth
    from betse_b200.engine import TissueEngine
    from betse_b200.partition import gather, partition
    from betse_b200.strips import DistributedStrips
    fields = ["cc_cells", "cc_env", "vm", "gjopen", "E_env_x", "v_env"]
    mesh, p, st = synth.make_tissue(args.cells)
    ds = DistributedStrips(mesh, p, st, local_rank, dist)
    ds.update_V()
    status = ds.step(args.steps)
    mine = ds.download_local(fields)
    allf = [None] * dist.get_world_size()
    dist.all_gather_object(allf, mine)
    ds.close()
    ok = True
    if rank == 0:
        parts = partition(mesh, p, st, dist.get_world_size())
        got = gather(parts, allf)
        eng = TissueEngine(mesh, p, st, device=local_rank)
        eng.update_V()
        s0 = eng.step(args.steps)
        ref = eng.download(fields)
        eng.close()
        worst = {}
        for f in fields:
            a, b = got[f].reshape(ref[f].shape), ref[f]
            same = bool(np.array_equal(a, b))
            worst[f] = 0.0 if same else float(np.max(np.abs(a - b)))
            ok &= same
        print(json.dumps({"check": "strips == single domain (bit-exact)", "ok": ok, "world": dist.get_world_size(),
                          "cells": len(mesh["cell_vol"]), "steps": args.steps, "status": [int(status), int(s0)],
                          "max_abs_diff": worst}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
