#!/bin/bash
mkdir -p gpurun_out
echo "== streamk pdl + prefetch"; timeout 300 python tools/exp_streamk.py --pdl --prefetch --m 1,4 --shapes up,gateup,down,o,qkv,gateup70 2>&1 | tee gpurun_out/r26_streamk_pf.log
