#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell" -s 4 -c 1 -o gpurun_out/r02x_kcell python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02x_ncu.log 2>&1
echo "ncu rc=$?"
