#!/bin/bash
for v in NOMMA NOEXTRACT; do echo "== $v"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_$v.so timeout 300 python tools/exp_streamk.py --m 1 --shapes gateup70 2>&1 | tail -1; done
echo "== base"; timeout 300 python tools/exp_streamk.py --m 1 --shapes gateup70 2>&1 | tail -1
