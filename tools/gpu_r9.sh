#!/bin/bash
mkdir -p gpurun_out
for v in occ2 occ4; do echo "== $v"; Q100_LIB=$PWD/llama3-quantization_b200/q100/libq100_$v.so timeout 300 python tools/exp_streamk.py --m 1 --shapes up,gateup,gateup70 2>&1 | tail -3; done
timeout 300 python tools/exp_streamk.py --m 1 --shapes gateup70 > gpurun_out/plain_sk.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k3_skinny' -s 20 -c 1 -o gpurun_out/prof_r1e_skinny python tools/exp_streamk.py --m 1 --shapes gateup70 > gpurun_out/ncu_sk.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/ncu_sk.log; cat gpurun_out/plain_sk.log
