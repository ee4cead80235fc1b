#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_block.py tests/test_block_kernels.py tests/test_parity_gemm_tc.py -m gpu -x -q --timeout 300 > gpurun_out/t_blk.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/t_blk.log
export Q100_BENCH_SKIP_MODEL=1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r55_bench.log 2> gpurun_out/r55_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r55_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r55_bench.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1))
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4), 'api', round(v['api_7_launches_ms_per_block'],4))
print({k:round(v['ms'],4) for k,v in d['quantize'].items()})
PY
