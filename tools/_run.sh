set -x
python tools/profile_r02.py i8 > /dev/null 2>&1 && python tools/profile_r02.py attn > /dev/null 2>&1 || echo "plain run failed"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k3_i8 -s 2 -c 1 -f -o gpurun_out/r02c_i8 python tools/profile_r02.py i8 > gpurun_out/ncu_i8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k8_attn -s 2 -c 1 -f -o gpurun_out/r02c_attn_b16 python tools/profile_r02.py attn > gpurun_out/ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k8_attn -s 5 -c 1 -f -o gpurun_out/r02c_attn_b1 python tools/profile_r02.py attn > gpurun_out/ncu_attn1.log 2>&1
for f in r02c_i8 r02c_attn_b16 r02c_attn_b1; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null; done
Q100_BENCH_SKIP_MODEL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02c_launches_raw.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/r02c_*
