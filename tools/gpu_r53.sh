#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_full.sh
cp gpurun_out/bench.log gpurun_out/r01l_bench.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 4 -c 1 -o gpurun_out/r01l_skinny_m8 python tools/prof_decode_m.py 8 > gpurun_out/r01l_ncu_m8.log 2>&1; echo "m8 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 4 -c 1 -o gpurun_out/r01l_skinny_m1 python tools/prof_decode_m.py 1 > gpurun_out/r01l_ncu_m1.log 2>&1; echo "m1 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'k2_rows' -s 2 -c 1 -o gpurun_out/r01l_k2 python tools/profile_kernels.py --iters 1 --which k2 --dtype f16 > gpurun_out/r01l_ncu_k2.log 2>&1; echo "k2 rc=$?"
for n in skinny_m8 skinny_m1 k2; do ncu -i gpurun_out/r01l_$n.ncu-rep --page raw --csv > gpurun_out/r01l_${n}_raw.csv 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
