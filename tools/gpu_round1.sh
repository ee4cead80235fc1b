#!/bin/bash
# First GPU call: smoke, parity tests, bench, launch list.  Everything logs to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== parity quant"; timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -q -x --timeout 300 > gpurun_out/t_quant.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/t_quant.log
echo "== parity gemm (skinny)"; timeout 900 python -m pytest tests/test_parity_gemm.py -m gpu -q --timeout 300 > gpurun_out/t_gemm.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/t_gemm.log
echo "== parity gemm (tc)"; timeout 600 python -m pytest tests/test_parity_gemm_tc.py -m gpu -q --timeout 200 > gpurun_out/t_tc.log 2>&1; echo "rc=$?"; tail -30 gpurun_out/t_tc.log
echo "== kernels"; timeout 300 python tools/profile_kernels.py --iters 20 > gpurun_out/kernels.log 2>&1; echo "rc=$?"; cat gpurun_out/kernels.log | tail -12
echo "== kernels (no tc)"; timeout 300 python tools/profile_kernels.py --iters 20 --which k1,k1pack,k2,skinny1,skinny16 > gpurun_out/kernels_notc.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/kernels_notc.log
echo "== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; echo "rc=$?"; tail -c 3000 gpurun_out/bench.log
