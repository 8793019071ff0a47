#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 900 > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02d_pytest.log
tail -8 gpurun_out/r02d_pytest.log
for c in c1 c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 200 --warmup 20 --cpu-budget 10 > gpurun_out/r02d_bench_$c.json 2> gpurun_out/r02d_bench_$c.err
  echo "$c rc=$?"; tail -c 600 gpurun_out/r02d_bench_$c.err
done
for b in 16 64; do
  timeout 600 python bench.py --config c2 --ensemble $b --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02d_bench_c2_ens$b.json 2> gpurun_out/r02d_bench_c2_ens$b.err
  echo "c2 ens $b rc=$?"; tail -c 600 gpurun_out/r02d_bench_c2_ens$b.err
done
timeout 600 python bench.py --config c1 --ensemble 128 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/r02d_bench_c1_ens128.json 2> gpurun_out/r02d_bench_c1_ens128.err
echo "c1 ens rc=$?"; tail -c 600 gpurun_out/r02d_bench_c1_ens128.err
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/r02d_bench_c5_20.json 2> gpurun_out/r02d_bench_c5_20.err
echo "c5 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02d_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'ms/step %.4f value %.3e e2e %s roof %.3f step %.3f' % (d['ms_per_step'], d['value'], d.get('e2e',{}).get('value'), d['roofline']['frac'], d['roofline']['step']['frac']), d['roofline']['kernel_ms'], d.get('ensemble'), d.get('cpu_baseline',{}).get('value'))
    except Exception as e:
        print(f, 'ERR', e)
PY
