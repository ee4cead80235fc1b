#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/profile_kernels.py --iters 1 --which skinny1 > gpurun_out/plain_sk.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 2 -c 1 -o gpurun_out/prof_r1c_skinny python tools/profile_kernels.py --iters 1 --which skinny1 > gpurun_out/ncu_sk.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_sk.log
