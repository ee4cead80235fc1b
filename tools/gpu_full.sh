#!/bin/bash
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/t_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench"; timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'ms',round(d['ms_per_step'],3),'roofline',round(d['roofline']['frac'],3),'e2e',round(d['e2e']['value'],1))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']),'frac',round(v['frac_hbm'],3),'tok/s',round(v['linear_tokens_per_s_32_blocks']), v['per_launch'], 'api', round(v['api_7_launches_ms_per_block'],3))
print(d['quantize']); print(d['clocks'])
PY
