#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_quant.py tests/test_block.py tests/test_parity_gemm.py -m gpu -x -q --timeout 300 > gpurun_out/t_q.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_q.log
timeout 300 python tools/profile_kernels.py --iters 20 --which k2 --dtype f16 --graph
export Q100_BENCH_SKIP_MODEL=1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r50_bench.log 2> gpurun_out/r50_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r50_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r50_bench.log').read().strip().splitlines()[-1])
print('value',round(d['value'],1),'e2e',round(d['e2e']['value'],1)); print(d['quantize']); print(d['roofline'])
PY
