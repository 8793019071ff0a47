#!/bin/bash
# usage (on the GPU box): tools/sweep.sh  -> prints ms/step and kernel times for env-var variants
for ov in 1 0; do
  BETSE_OVERLAP=$ov python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap=$ov', round(d['ms_per_step'],4), {k:round(v,4) for k,v in d['roofline']['kernel_ms'].items()})"
done
