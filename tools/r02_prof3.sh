#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_chan_cell|k_chan_env_cell|k_cell_update|k_field" -s 8 -c 4 -o gpurun_out/r02s_chan python bench.py --config c3 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02s_ncu_full.log 2>&1
echo "ncu rc=$?"
