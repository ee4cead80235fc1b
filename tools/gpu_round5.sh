#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 -x > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gemm.log
echo "== streamk exp"; timeout 300 python tools/exp_streamk.py 2>&1 | tail -9
echo "== kernels (graph)"; timeout 300 python tools/profile_kernels.py --iters 12 --graph --which skinny1,skinny16 2>&1 | tail -3
