#!/bin/bash
# re-entry: full GPU test pass + bench + round-1 profiles (launch list, ncu --set full of the headline kernels)
mkdir -p gpurun_out
bash tools/gpu_full.sh
bash tools/gpu_prof_r1h.sh
cp gpurun_out/bench.log gpurun_out/r1h_bench.json
