#!/bin/bash
mkdir -p gpurun_out
echo "== parity quant"; timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -q --timeout 300 -x > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_quant.log
echo "== kernels"; timeout 300 python tools/profile_kernels.py --iters 12 --graph --which k1,k1pack,k2 --dtype f16 2>&1 | tail -4
timeout 300 python tools/profile_kernels.py --iters 12 --graph --which k1,k1pack --dtype bf16 2>&1 | tail -3
