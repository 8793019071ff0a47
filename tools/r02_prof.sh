#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell|k_envacc_ell|k_ion|k_field" -s 24 -c 6 -o gpurun_out/r02k_step python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02k_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02k_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02k_launches.log 2>&1
echo "launches rc=$?"
