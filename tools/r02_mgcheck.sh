#!/bin/bash
N=${1:-4}; CELLS=${2:-30000}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
for v in "BETSE_XFUSE=1" "BETSE_XFUSE=1 BETSE_OVERLAP=0"; do
  for rep in 1 2 3 4 5; do
    echo -n "$v : "; env $v timeout 200 $TR tools/check_multigpu.py --cells $CELLS --steps 40 2>/dev/null | tail -1 | cut -c1-260
  done
done
