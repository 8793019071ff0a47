#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell_patch|k_envacc_patch" -s 4 -c 2 -o gpurun_out/r02l_patch python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02l_ncu.log 2>&1
echo "ncu rc=$?"
