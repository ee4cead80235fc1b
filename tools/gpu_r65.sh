#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -x -q --timeout 300 > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/t_quant.log
for dt in f16 bf16; do timeout 300 python tools/profile_kernels.py --iters 20 --which k1,k1pack --dtype $dt --graph; done
