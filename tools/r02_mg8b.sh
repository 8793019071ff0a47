#!/bin/bash
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/check_multigpu.py --cells 60000 --steps 40 2>/dev/null | tail -1 | cut -c1-200
timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02y_bench8.json 2> gpurun_out/r02y_bench8.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench8.json').read().strip().splitlines()[-1])
print('N=8 ms/step %.4f value %.3e e2e %.3e' % (d['ms_per_step'], d['value'], d['e2e']['value']), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, d['strips_parity']['ok'], d['e2e'].get('seconds'))
PY
BETSE_FIELD16=0 timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 20 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('FIELD16=0', round(d['ms_per_step'],4), {k:round(x,4) for k,x in d['roofline']['kernel_ms'].items()})"
