#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gemm.log
echo "== parity gemm (tc)"; timeout 600 python -m pytest tests/test_parity_gemm_tc.py -m gpu -q --timeout 200 > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_tc.log
echo "== parity quant"; timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -q -x --timeout 300 > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_quant.log
echo "== kernels (graph)"; timeout 300 python tools/profile_kernels.py --iters 12 --graph > gpurun_out/kernels_graph.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/kernels_graph.log
