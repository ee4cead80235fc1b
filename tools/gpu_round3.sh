#!/bin/bash
mkdir -p gpurun_out
echo "== parity quant"; timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -q -x --timeout 300 > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_quant.log
echo "== parity gemm (tc)"; timeout 600 python -m pytest tests/test_parity_gemm_tc.py -m gpu -q --timeout 200 > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_tc.log
echo "== kernels"; timeout 300 python tools/profile_kernels.py --iters 20 --which k1,k1pack,tc > gpurun_out/kernels.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/kernels.log
echo "== ncu tc"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_tc' -s 3 -c 1 -o gpurun_out/prof_r1b_tc python tools/profile_kernels.py --iters 1 --which tc > gpurun_out/ncu_tc.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_tc.log
