#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -5
for c in c2 c3 c4; do timeout 600 python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()})"; done
tools/sweep.sh "BETSE_X=1"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02t_launches_c3.csv python bench.py --config c3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02t_ncu_c3.log 2>&1
