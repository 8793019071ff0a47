#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_channels.py tests/test_gpu_network.py tests/test_full_run.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 900 2>&1 | tail -5
for c in c3 c4; do timeout 600 python bench.py --config $c --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$c', round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()})"; done
tools/sweep.sh "BETSE_X=1"
timeout 600 python bench.py --steps 200 --warmup 20 --e2e-kind SIM --no-cpu-baseline > gpurun_out/r02q_bench_c5_200_sim.json 2>/dev/null; python -c "
import json; d=json.loads(open('gpurun_out/r02q_bench_c5_200_sim.json').read().strip().splitlines()[-1]); print('SIM e2e %.3e'%d['e2e']['value'], d['e2e'].get('seconds'))"
