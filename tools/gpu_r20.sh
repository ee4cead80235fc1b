#!/bin/bash
for v in NOEPI; do echo "== $v"; Q100_TCA_CLUSTER=1 Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_$v.so timeout 300 python tools/profile_kernels.py --iters 12 --graph --which tc 2>&1 | tail -1; done
