#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.log
timeout 300 python tools/profile_kernels.py --iters 3 --which tc > gpurun_out/plain_tc.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_tca' -s 3 -c 1 -o gpurun_out/prof_r1g_tca python tools/profile_kernels.py --iters 3 --which tc > gpurun_out/ncu_tc.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/ncu_tc.log
