#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_block.py tests/test_block_kernels.py -m gpu -x -q --timeout 300 > gpurun_out/t_blk.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_blk.log
timeout 900 python tools/bench_model.py --layers 32 --ppl-samples 0 --batches 1,4,16 > gpurun_out/r49_model.log 2>&1; echo "model rc=$?"; tail -1 gpurun_out/r49_model.log | cut -c1-900
