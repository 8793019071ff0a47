#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_synth.py tests/test_gpu_strips.py tests/test_gpu_ensemble.py -m gpu -q -x --timeout 600 > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02f_pytest.log
tail -15 gpurun_out/r02f_pytest.log
timeout 900 tools/sweep.sh "BETSE_X=1" "BETSE_ENVDEPS=0" "BETSE_KCELL_SHARE=0 BETSE_ENVDEPS=0" "BETSE_KCELL_SHARE=0" "BETSE_OVERLAP=0" > gpurun_out/r02f_sweep.txt 2>&1
cat gpurun_out/r02f_sweep.txt
