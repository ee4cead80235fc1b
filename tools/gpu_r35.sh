#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -x -q --timeout 300 > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_quant.log
for dt in f16 bf16; do timeout 300 python tools/profile_kernels.py --iters 20 --which k1,k1pack --dtype $dt --graph; done 2>&1 | tee gpurun_out/r35_k1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_fast' -c 6 -o gpurun_out/r35_k1fast python tools/profile_kernels.py --iters 1 --which k1,k1pack --dtype f16 > gpurun_out/r35_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r35_k1fast.ncu-rep --page raw --csv > gpurun_out/r35_k1fast_raw.csv 2>/dev/null
