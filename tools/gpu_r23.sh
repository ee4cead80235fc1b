#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_block.py -m gpu -x -q --timeout 300 2>&1 | tail -15
