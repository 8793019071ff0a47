#!/bin/bash
N=${1:-2}
TAG=${2:-r02i}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/check_multigpu.py --cells 60000 --steps 12 2>gpurun_out/${TAG}_check$N.err | tail -1
timeout 300 $TR tools/check_multigpu.py --cells 1000000 --steps 25 2>>gpurun_out/${TAG}_check$N.err | tail -1
for v in ${VARIANTS:-"BETSE_XFUSE=1" "BETSE_XFUSE=0" "BETSE_XWAIT=0"}; do
  env $v timeout 600 $TR bench.py --gpus $N --steps 200 --warmup 20 --no-e2e > gpurun_out/${TAG}_bench${N}_$v.json 2> gpurun_out/${TAG}_bench${N}_$v.err
  echo "$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${TAG}_bench${N}_$v.json').read().strip().splitlines()[-1])
    print('N=$N $v ms/step %.4f value %.3e' % (d['ms_per_step'], d['value']), {k: round(x,4) for k,x in d['roofline']['kernel_ms'].items()}, 'finite', d['finite'], 'status', d['status_word'])
except Exception as e:
    print('ERR', e); print(open('gpurun_out/${TAG}_bench${N}_$v.err').read()[-1500:])
PY
done
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/${TAG}_check$N.err | tail -5
