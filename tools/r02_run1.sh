#!/bin/bash
# first GPU pass of round 2: parity suite, A/B of the membrane kernels, ncu of k_cell
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
tail -5 gpurun_out/r02a_pytest.log
tools/sweep.sh "BETSE_KCELL=0" "BETSE_KCELL=1 BETSE_KCELL_MINB=2" "BETSE_KCELL=1 BETSE_KCELL_MINB=3" "BETSE_KCELL=1 BETSE_KCELL_MINB=4" > gpurun_out/r02a_sweep.txt 2>&1
cat gpurun_out/r02a_sweep.txt
BETSE_KCELL_MINB=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell|k_envacc_ell" -s 6 -c 2 -o gpurun_out/r02a_kcell python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02a_ncu.log 2>&1
echo "ncu rc=$?"
