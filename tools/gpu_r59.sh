#!/bin/bash
# SURVEY 8d configs 3 and 4 on one GPU: W2 / W3 / W4 A16 g128 and W8A8 / W4A4 (per-token activation fake-quant), 32 blocks,
# perplexity against the torch port on the same weights + prefill / decode rates
mkdir -p gpurun_out
: > gpurun_out/r01l_model_configs.jsonl
for cfg in "--wbits 2" "--wbits 3" "--wbits 4" "--wbits 8 --abits 8" "--wbits 4 --abits 4"; do
  echo "== $cfg"
  timeout 500 python tools/bench_model.py --layers 32 --ppl-samples 1 --ppl-ref --batches 1,16 $cfg 2>gpurun_out/m.err | tail -1 | tee -a gpurun_out/r01l_model_configs.jsonl | cut -c1-700
  tail -1 gpurun_out/m.err | cut -c1-200
done
