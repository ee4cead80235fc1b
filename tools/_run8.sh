timeout 840 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r4b_bench_n8.json 2> gpurun_out/r4b_bench_n8.err
echo rc=$?
tail -c 1500 gpurun_out/r4b_bench_n8.err
wc -c gpurun_out/r4b_bench_n8.json
