#!/bin/bash
for v in NO_ONES NO_MMA; do echo "== $v"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_$v.so timeout 300 python tools/exp_streamk.py --m 1,16 --shapes up,gateup70 2>&1 | tail -4; done
