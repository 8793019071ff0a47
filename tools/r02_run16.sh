#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_golden.py tests/test_gpu_dropin.py -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 300 python tools/time_hh.py
BETSE_HH_SYM=0 timeout 300 python tools/time_hh.py
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_hh" -c 40 --csv --log-file gpurun_out/r02v_hh_launches.csv python tools/time_hh.py > gpurun_out/r02v_hh.log 2>&1
