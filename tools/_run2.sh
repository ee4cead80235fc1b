timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/reduce_norm_check.py > gpurun_out/rn.log 2>&1
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/rn.log | grep -A12 "rank0.*Traceback\|world\|OK" | head -30
