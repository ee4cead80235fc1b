timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/reduce_norm_check.py > gpurun_out/rn8.log 2>&1
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/rn8.log | grep -A12 "rank0.*Traceback\|world\|OK" | head -20
