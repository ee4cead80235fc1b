#!/bin/bash
# round-1 final profiles: launch list of the bench command + ncu --set full of the four headline kernels
mkdir -p gpurun_out
export Q100_BENCH_SKIP_MODEL=1
timeout 600 python bench.py --steps 2 --warmup 3 > gpurun_out/r1h_bench_plain.log 2> gpurun_out/r1h_bench_plain.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1h_launches_raw.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r1h_ncu_launch.log 2>&1; echo "launch-list rc=$?"
timeout 300 python tools/profile_kernels.py --iters 1 --which k1,k1pack,skinny1,tc > gpurun_out/r1h_plain.log 2>&1; echo "plain rc=$?"; cat gpurun_out/r1h_plain.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k3_tca|k3_skinny|k1_groups' -s 8 -c 4 -o gpurun_out/r1h_full python tools/profile_kernels.py --iters 1 --which k1,k1pack,skinny1,tc > gpurun_out/r1h_ncu_full.log 2>&1; echo "full rc=$?"; tail -3 gpurun_out/r1h_ncu_full.log
ncu -i gpurun_out/r1h_full.ncu-rep --page raw --csv > gpurun_out/r1h_full_raw.csv 2>/dev/null; wc -c gpurun_out/r1h_full_raw.csv
