#!/bin/bash
Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_TIMING.so timeout 300 python tools/exp_timing.py 2>&1 | tail -40
