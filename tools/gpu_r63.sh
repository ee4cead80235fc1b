#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gemm.py tests/test_block.py tests/test_block_kernels.py -m gpu -x -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_gemm.log
timeout 500 python tools/bench_model.py --layers 32 --ppl-samples 0 --batches 1,4,16 2>gpurun_out/m.err | tail -1 | cut -c1-900
