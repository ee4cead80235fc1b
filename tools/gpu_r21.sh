#!/bin/bash
echo "== default"; timeout 300 python tools/exp_tc_shapes.py 2>&1 | tail -5
echo "== NT=256 forced"; Q100_TCA_NT=256 timeout 300 python tools/exp_tc_shapes.py 2>&1 | tail -5
echo "== NT=128 forced"; Q100_TCA_NT=128 timeout 300 python tools/exp_tc_shapes.py 2>&1 | tail -5
