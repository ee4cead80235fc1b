#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader
echo "== micro hmma"; timeout 60 tools/micro/hmma_bench 2>&1 | tail -12
echo "== micro stream2"; timeout 120 tools/micro/stream_bench2 2>&1 | tail -24
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 -x > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -12 gpurun_out/t_gemm.log
echo "== sweep minb4 bf16"; timeout 300 python tools/exp_streamk.py --m 1,16 2>&1 | tail -14
echo "== sweep minb4 bf16 pdl"; timeout 300 python tools/exp_streamk.py --m 1 --pdl 2>&1 | tail -14
echo "== sweep minb4 f16"; timeout 300 python tools/exp_streamk.py --m 1 --dtype f16 --shapes up,gateup,gateup70 2>&1 | tail -14
echo "== sweep minb3 bf16"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_minb3.so timeout 300 python tools/exp_streamk.py --m 1 2>&1 | tail -14
