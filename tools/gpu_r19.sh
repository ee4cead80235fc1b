#!/bin/bash
mkdir -p gpurun_out
echo "== parity tc"; timeout 600 python -m pytest tests/test_parity_gemm_tc.py -m gpu -q --timeout 120 -x > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/t_tc.log
for c in 1 2 4; do echo "== cluster $c"; Q100_TCA_CLUSTER=$c timeout 300 python tools/profile_kernels.py --iters 12 --graph --which tc 2>&1 | tail -1; done
