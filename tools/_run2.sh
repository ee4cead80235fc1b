timeout 1100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r4f_bench_n2.json 2> gpurun_out/r4f_bench_n2.err
echo rc=$?
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r4f_bench_n2.err | tail -8
wc -c gpurun_out/r4f_bench_n2.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 tools/tp_parity_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -12
