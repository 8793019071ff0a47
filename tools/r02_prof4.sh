#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_hh_gemm" -s 4 -c 1 -o gpurun_out/r02u_gemm python tools/time_hh.py > gpurun_out/r02u_ncu_full.log 2>&1
echo "ncu rc=$?"
