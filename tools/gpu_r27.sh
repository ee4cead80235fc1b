#!/bin/bash
Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_TIMING.so timeout 200 python tools/exp_walltime.py 2>&1 | tee gpurun_out/r27_wall.log
