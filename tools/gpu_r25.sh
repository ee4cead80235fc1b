#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -x -q --timeout 300 2>&1 | tail -5
echo "== streamk pdl"; timeout 300 python tools/exp_streamk.py --pdl --m 1,4,16 --shapes up,gateup,down,o,qkv,gateup70 2>&1 | tee gpurun_out/r25_streamk.log
