#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_golden.py tests/test_gpu_synth.py -m gpu -q -x --timeout 900 > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02l_pytest.log
tail -12 gpurun_out/r02l_pytest.log
timeout 900 tools/sweep.sh "BETSE_X=1" "BETSE_OVERLAP=0" > gpurun_out/r02l_sweep.txt 2>&1
cat gpurun_out/r02l_sweep.txt
