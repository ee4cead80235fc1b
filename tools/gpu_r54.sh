#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gemm.py tests/test_block.py -m gpu -x -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_gemm.log
timeout 600 python tools/exp_decode_paths.py 2>&1 | cut -c1-110 | tee gpurun_out/r54_paths.log
