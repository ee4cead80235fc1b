#!/bin/bash
mkdir -p gpurun_out
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 -x > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gemm.log
echo "== sweep bf16"; timeout 300 python tools/exp_streamk.py --m 1,16 2>&1 | tail -14
echo "== sweep bf16 pdl"; timeout 300 python tools/exp_streamk.py --m 1 --pdl 2>&1 | tail -14
echo "== sweep f16"; timeout 300 python tools/exp_streamk.py --m 1,16 --dtype f16 --shapes up,gateup,gateup70 2>&1 | tail -14
