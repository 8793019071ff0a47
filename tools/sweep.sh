#!/bin/bash
# usage (on the GPU box): tools/sweep.sh "VAR=val VAR2=val" "VAR=val2" ...  -> ms/step and kernel times per env-var variant
for v in "$@"; do
  env $v python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'finite', d.get('finite'), 'status', d.get('status_word'))"
done
