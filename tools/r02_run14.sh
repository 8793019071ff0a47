#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 900 2>&1 | tail -5
timeout 300 python tools/time_hh.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hh|k_diag" -c 60 --csv --log-file gpurun_out/r02u_hh_launches.csv python tools/time_hh.py > gpurun_out/r02u_hh.log 2>&1
