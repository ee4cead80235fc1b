#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 120 -x > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_gemm.log
echo "== sweep bf16"; timeout 300 python tools/exp_streamk.py --m 1,8,16 2>&1 | tail -18
echo "== sweep bf16 pdl"; timeout 300 python tools/exp_streamk.py --m 1 --pdl 2>&1 | tail -14
