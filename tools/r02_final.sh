#!/bin/bash
mkdir -p gpurun_out
T=r02z
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/${T}_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
tail -6 gpurun_out/${T}_pytest.log
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/${T}_bench_c5_20.json 2> gpurun_out/${T}_bench_c5_20.err; echo "c5/20 rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --cpu-budget 10 > gpurun_out/${T}_bench_c5_200.json 2> gpurun_out/${T}_bench_c5_200.err; echo "c5/200 rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --e2e-kind SIM --no-cpu-baseline > gpurun_out/${T}_bench_c5_200_sim.json 2> /dev/null; echo "c5/sim rc=$?"
timeout 600 python bench.py --steps 200 --warmup 20 --ragged 0.2 --no-cpu-baseline > gpurun_out/${T}_bench_c5_ragged.json 2> gpurun_out/${T}_bench_c5_ragged.err; echo "ragged rc=$?"
for c in c1 c2 c3 c4; do
  timeout 600 python bench.py --config $c --steps 200 --warmup 20 --cpu-budget 8 > gpurun_out/${T}_bench_$c.json 2> gpurun_out/${T}_bench_$c.err; echo "$c rc=$?"
done
timeout 600 python bench.py --config c2 --ensemble 64 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/${T}_bench_c2_ens64.json 2>/dev/null; echo "ens c2 rc=$?"
timeout 600 python bench.py --config c1 --ensemble 128 --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/${T}_bench_c1_ens128.json 2>/dev/null; echo "ens c1 rc=$?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${T}_bench_reference_arm.json 2>/dev/null; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r02z_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'ms/step %.4f value %.3e e2e %.3e' % (d['ms_per_step'], d['value'], d.get('e2e',{}).get('value',0)), 'roof %.3f step %.3f' % (d.get('roofline',{}).get('frac',0), d.get('roofline',{}).get('step',{}).get('frac',0)), {k: round(x,4) for k,x in d.get('roofline',{}).get('kernel_ms',{}).items()}, 'cpu', d.get('cpu_baseline',{}).get('value'), d.get('ensemble',{}).get('speedup_vs_solo_per_gpu'))
    except Exception as e:
        print(f, 'ERR', e)
PY
