#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_quant.py -m gpu -x -q --timeout 300 -k "backward" > gpurun_out/t_bwd.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/t_bwd.log
# launch list of the decode step of a 4-block model (graph replays are traced per kernel by ncu)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node -c 3000 --csv --log-file gpurun_out/r48_decode_launches_raw.csv python tools/bench_model.py --layers 4 --ppl-samples 0 --batches 1 --decode-steps 6 > gpurun_out/r48_model.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r48_model.log | cut -c1-400
