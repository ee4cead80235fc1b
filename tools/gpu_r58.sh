#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_quant.py tests/test_block.py -m gpu -x -q --timeout 300 > gpurun_out/t_q.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_q.log
timeout 300 python tools/profile_kernels.py --iters 20 --which k2 --dtype f16 --graph
timeout 300 python tools/profile_kernels.py --iters 20 --which k2 --dtype bf16 --graph
