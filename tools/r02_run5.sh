#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_synth.py tests/test_gpu_strips.py -m gpu -q -x --timeout 600 > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02e_pytest.log
tail -15 gpurun_out/r02e_pytest.log
tools/sweep.sh "BETSE_KCELL_PIPE=1" "BETSE_KCELL_PIPE=0" "BETSE_KCELL_PIPE=0 BETSE_ENVACC_REV=1" "BETSE_KCELL_PIPE=1 BETSE_ENVACC_REV=1" > gpurun_out/r02e_sweep.txt 2>&1
cat gpurun_out/r02e_sweep.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell" -s 4 -c 1 -o gpurun_out/r02e_kcell_pipe python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02e_ncu.log 2>&1
echo "ncu rc=$?"
