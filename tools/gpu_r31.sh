#!/bin/bash
timeout 900 python tools/bench_model.py --ppl-samples 0 > gpurun_out/r31_model.log 2> gpurun_out/r31_model.err; echo "rc=$?"; tail -3 gpurun_out/r31_model.err; cat gpurun_out/r31_model.log
