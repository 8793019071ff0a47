#!/bin/bash
mkdir -p gpurun_out
for c in c3 c4; do
BETSE_CHAN_PASS=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02r_launches_$c.csv python bench.py --config $c --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02r_ncu_$c.log 2>&1
done
tail -3 gpurun_out/r02r_ncu_c3.log
