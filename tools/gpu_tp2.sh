#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_tp2.log 2> gpurun_out/bench_tp2.err; echo "rc=$?"; tail -5 gpurun_out/bench_tp2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tp2.log').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'roofline',round(d['roofline']['frac'],3))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']),'frac',round(v['frac_hbm'],3),'graph',v['cuda_graph'])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
echo "== TP2 model: perplexity vs the torch port, prefill, decode"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/bench_model.py --layers 4 --ppl-samples 1 --ppl-ref --batches 1,4 2>&1 | tail -1 | cut -c1-900
