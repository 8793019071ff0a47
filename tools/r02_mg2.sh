#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR tools/check_dropin_multigpu.py > gpurun_out/r02m_dropin2.json 2> gpurun_out/r02m_dropin2.err
echo "dropin rc=$?"; cat gpurun_out/r02m_dropin2.json; grep -v "OMP_NUM\|\*\*\*\*\|Warning\|warn" gpurun_out/r02m_dropin2.err | tail -15
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/check_multigpu.py --cells 60000 --steps 12 2>/dev/null | tail -1
timeout 600 $TR bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/r02m_bench2.json 2> gpurun_out/r02m_bench2.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r02m_bench2.json
