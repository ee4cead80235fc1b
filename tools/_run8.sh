export Q100_PEER_VERBOSE=1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/peer_scatter_check.py > gpurun_out/mc8.log 2>&1
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/mc8.log | grep -A8 "rank0.*Traceback\|world\|OK" | head -30
Q100_NO_MULTIMEM=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tools/peer_scatter_check.py > gpurun_out/sc8.log 2>&1
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/sc8.log | grep -A8 "rank0.*Traceback\|world\|OK" | head -30
