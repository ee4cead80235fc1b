#!/bin/bash
timeout 900 python -m pytest tests/test_block_kernels.py tests/test_block.py -m gpu -x -q --timeout 300 2>&1 | tail -15
timeout 900 python tools/bench_model.py --ppl-samples 0 > gpurun_out/r30_model.log 2> gpurun_out/r30_model.err; echo "rc=$?"; tail -3 gpurun_out/r30_model.err; cat gpurun_out/r30_model.log
