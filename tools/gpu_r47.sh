#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -x -q --timeout 300 -k "backward" > gpurun_out/t_bwd.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_bwd.log
timeout 300 python tools/exp_backward.py 2>&1 | tee gpurun_out/r47_bwd.log
export Q100_BENCH_SKIP_MODEL=1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r47_bench.log 2> gpurun_out/r47_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r47_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r47_bench.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',d['e2e'])
PY
