#!/bin/bash
mkdir -p gpurun_out
echo "== stream_bench3"; timeout 120 tools/micro/stream_bench3 2>&1 | tee gpurun_out/r22_stream3.log
echo "== streamk pdl"; timeout 300 python tools/exp_streamk.py --pdl --m 1,4,16 2>&1 | tee gpurun_out/r22_streamk.log
echo "== k1"; timeout 120 python tools/exp_k1.py 2>&1 | tee gpurun_out/r22_k1.log
echo "== sk timing"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_TIMING.so timeout 200 python tools/exp_timing.py 2>&1 | tee gpurun_out/r22_timing.log
