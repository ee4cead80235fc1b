#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/bench_model.py --ppl-ref > gpurun_out/r24_model.log 2> gpurun_out/r24_model.err; echo "rc=$?"; tail -3 gpurun_out/r24_model.err; cat gpurun_out/r24_model.log
