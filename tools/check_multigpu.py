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
    from betse_b200.strips import verify_strips
    out = verify_strips(dist, local_rank, cells=args.cells, steps=args.steps)
    ok = True
    if rank == 0:
        ok = out["ok"]
        print(json.dumps(out))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
