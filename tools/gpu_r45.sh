#!/bin/bash
# full round-1 validation + profiles r01k
mkdir -p gpurun_out
bash tools/gpu_full.sh
cp gpurun_out/bench.log gpurun_out/r01k_bench.json
export Q100_BENCH_SKIP_MODEL=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r01k_launches_raw.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r01k_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_tca' -s 2 -c 1 -o gpurun_out/r01k_tca python tools/profile_kernels.py --iters 1 --which tc > gpurun_out/r01k_ncu_tca.log 2>&1; echo "tca rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 2 -c 1 -o gpurun_out/r01k_skinny python tools/profile_kernels.py --iters 1 --which skinny1 > gpurun_out/r01k_ncu_skinny.log 2>&1; echo "skinny rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_fast' -s 2 -c 1 -o gpurun_out/r01k_k1fake python tools/profile_kernels.py --iters 1 --which k1 --dtype f16 > gpurun_out/r01k_ncu_k1.log 2>&1; echo "k1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_fast' -s 2 -c 1 -o gpurun_out/r01k_k1pack python tools/profile_kernels.py --iters 1 --which k1pack --dtype f16 > gpurun_out/r01k_ncu_k1p.log 2>&1; echo "k1pack rc=$?"
for n in tca skinny k1fake k1pack; do ncu -i gpurun_out/r01k_$n.ncu-rep --page raw --csv > gpurun_out/r01k_${n}_raw.csv 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
wc -c gpurun_out/r01k_*
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k1_backward' -s 2 -c 1 -o gpurun_out/r01k_k1b python tools/exp_backward.py > gpurun_out/r01k_ncu_k1b.log 2>&1; echo "k1b rc=$?"
ncu -i gpurun_out/r01k_k1b.ncu-rep --page raw --csv > gpurun_out/r01k_k1b_raw.csv 2>/dev/null; rm -f gpurun_out/*.ncu-rep
