#!/bin/bash
timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -x -q --timeout 300 2>&1 | tail -2
echo "== streamk NO pdl"; timeout 300 python tools/exp_streamk.py --m 1 --shapes up,gateup,down,o,qkv 2>&1 | tee gpurun_out/r29_streamk_nopdl.log
echo "== streamk pdl"; timeout 300 python tools/exp_streamk.py --pdl --m 1 --shapes up,gateup,down,o,qkv 2>&1 | tee gpurun_out/r29_streamk_pdl.log
Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_TIMING.so timeout 200 python tools/exp_walltime.py 2>&1 | tee gpurun_out/r29_wall.log
