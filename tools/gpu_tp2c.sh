#!/bin/bash
mkdir -p gpurun_out
for np in 0 1; do
  echo "== TP2 model, Q100_NO_PEER=$np"
  if [ $np = 1 ]; then export Q100_NO_PEER=1; else unset Q100_NO_PEER; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2952$np tools/bench_model.py --layers 8 --ppl-samples 1 --ppl-ref --batches 1,4,16 2>&1 | grep -v "OMP_NUM\|\*\*\*\*" | tail -1 | cut -c1-1000
done
unset Q100_NO_PEER
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_tp2.log 2> gpurun_out/bench_tp2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tp2.err | cut -c1-300
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_tp2.log').read().strip().splitlines()[-1])
print('n_gpus',d['n_gpus'],'value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],1))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']))
print(d['llama3_8b'])
PY
