#!/bin/bash
mkdir -p gpurun_out
python tools/build_ext.py > gpurun_out/build.log 2>&1
echo "== graph timing"; timeout 300 python tools/profile_kernels.py --iters 12 --graph > gpurun_out/kernels_graph.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/kernels_graph.log
echo "== ncu launch list"
timeout 300 python tools/profile_kernels.py --iters 3 > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python tools/profile_kernels.py --iters 3 > gpurun_out/ncu1.log 2>&1
echo "rc=$?"
echo "== ncu full (tc, skinny, k1)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k3_tc|k3_skinny|k1_groups' -c 9 -o gpurun_out/prof_r1a python tools/profile_kernels.py --iters 1 --which k1,skinny1,skinny16,tc > gpurun_out/ncu2.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/ncu2.log
ls -la gpurun_out
