for n in 4 8; do
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/r4g_bench_n$n.json 2> gpurun_out/r4g_bench_n$n.err
echo "n=$n rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r4g_bench_n$n.err | tail -5
wc -c gpurun_out/r4g_bench_n$n.json
done
