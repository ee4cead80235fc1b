#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r32.log 2> gpurun_out/bench_r32.err; echo "rc=$?"; tail -3 gpurun_out/bench_r32.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r32.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'roofline',round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value'],1))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']),'frac',round(v['frac_hbm'],3), v['per_launch'])
print(d['quantize']); print(d['llama3_8b']); print(d['cpu_baseline']); print(d['clocks'])
PY
