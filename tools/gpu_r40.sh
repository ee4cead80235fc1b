#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gemm.py tests/test_block.py tests/test_block_kernels.py -m gpu -x -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/t_gemm.log
for dt in bf16 f16; do timeout 300 python tools/profile_kernels.py --iters 20 --which skinny1 --dtype $dt --graph; done 2>&1 | tee gpurun_out/r40_skinny.log
export Q100_BENCH_SKIP_MODEL=1
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r40_bench.log 2> gpurun_out/r40_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r40_bench.log').read().strip().splitlines()[-1])
for k,v in d['decode'].items(): print(k, 'ms/block',round(v['ms_per_block'],4),'GB/s',round(v['GBps']),'frac',round(v['frac_hbm'],3), v['per_launch'])
print(d['quantize'])
PY
