#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/test_peer_allreduce.py > gpurun_out/peer.log 2>&1; echo "peer rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/peer.log | tail -12 | cut -c1-300
