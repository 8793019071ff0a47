#!/bin/bash
# usage (on the GPU box): tools/sweep_cells.sh "<cells> ..." "VAR=val" "VAR=val2" ... -> ms/step per tissue size and env-var variant
sizes=$1; shift
for n in $sizes; do for v in "$@"; do
  env $v python bench.py --cells $n --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$n', '$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()})"
done; done
