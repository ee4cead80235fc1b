#!/bin/bash
Q100_NVCC_FLAGS="-DQ100_EXP_NOCOMPUTE -DQ100_EXP_RR" python tools/build_ext.py --force > /dev/null 2>&1
python tools/exp_streamk.py 2>&1 | tail -9
