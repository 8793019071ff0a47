#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_channels.py tests/test_gpu_network.py tests/test_gpu_golden.py tests/test_full_run.py tests/test_gpu_ensemble.py -m gpu -q -x --timeout 900 2>&1 | tail -15
for v in 1 0; do BETSE_CHAN_CELL=$v timeout 600 python bench.py --config c3 --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('c3 cell=$v', round(d['ms_per_step'],4), '%.3e'%d['value'], 'e2e %.3e'%d['e2e']['value'], {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()})"; done
