#!/bin/bash
mkdir -p gpurun_out
N=${1:-4}
nvidia-smi -L | wc -l
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29530 tools/test_peer_allreduce.py > gpurun_out/peer$N.log 2>&1; echo "peer rc=$?"; grep -v "OMP_NUM\|\*\*\*\*\|File \|^  " gpurun_out/peer$N.log | head -8 | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_tp$N.log 2> gpurun_out/bench_tp$N.err; echo "bench rc=$?"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/bench_tp$N.err | tail -3 | cut -c1-300
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_tp$N.log').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']))
print(d['llama3_8b'])
PY
