#!/bin/bash
# usage (on the GPU box): tools/sweep.sh  -> prints ms/step and kernel times for env-var variants
for minb in 2 3 4; do for pf in 0 4096; do
  BETSE_KMEM_MINB=$minb BETSE_PF_TILES=$pf python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('minb=$minb pf=$pf', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"
done; done
