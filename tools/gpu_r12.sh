#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_gemm.log
Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_TIMING.so timeout 300 python tools/exp_timing.py 2>&1 | tail -36
echo "== sweep bf16"; timeout 300 python tools/exp_streamk.py --m 1,8,16 2>&1 | tail -18
echo "== sweep bf16 pdl"; timeout 300 python tools/exp_streamk.py --m 1 --pdl 2>&1 | tail -14
