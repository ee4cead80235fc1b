#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gemm.log
timeout 300 python tools/exp_streamk.py --m 1 --shapes gateup70 > gpurun_out/plain_sk.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 20 -c 1 -o gpurun_out/prof_r1d_skinny python tools/exp_streamk.py --m 1 --shapes gateup70 > gpurun_out/ncu_sk.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_sk.log; cat gpurun_out/plain_sk.log
