#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02c_pytest.log
tail -5 gpurun_out/r02c_pytest.log
tools/sweep.sh "BETSE_FUSE=0" "BETSE_FUSE=1" "BETSE_FUSE=0 BETSE_KCELL_STAGED=0" "BETSE_FUSE=1 BETSE_KCELL_STAGED=0" "BETSE_FUSE=1 BETSE_FUSE_LAG=600" "BETSE_FUSE=1 BETSE_FUSE_LAG=1200" > gpurun_out/r02c_sweep.txt 2>&1
cat gpurun_out/r02c_sweep.txt
BETSE_FUSE=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell|k_ion" -s 8 -c 2 -o gpurun_out/r02c_kcell python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02c_ncu.log 2>&1
echo "ncu rc=$?"
BETSE_FUSE=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_cell" -s 4 -c 1 -o gpurun_out/r02c_kcell_fused python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline >> gpurun_out/r02c_ncu.log 2>&1
echo "ncu rc=$?"
