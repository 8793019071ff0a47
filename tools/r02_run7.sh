#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02g_pytest.log
tail -5 gpurun_out/r02g_pytest.log
for rep in 1 2 3; do python tools/e2e_cprofile.py 20 2>&1 | grep -E "^rep"; done
for rep in 1 2; do python tools/e2e_cprofile.py 200 2>&1 | grep -E "^rep"; done
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02g_bench_c5_20.json 2> gpurun_out/r02g_bench_c5_20.err; echo "rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02g_bench_c5_200.json 2> gpurun_out/r02g_bench_c5_200.err; echo "rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02g_bench_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, 'ms/step %.4f value %.3e e2e %.3e' % (d['ms_per_step'], d['value'], d['e2e']['value']), d['e2e'].get('seconds'), d.get('cpu_baseline',{}).get('value'))
PY
