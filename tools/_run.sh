for cl in 2 1; do
echo "== CL=$cl"
Q100_I8_CLUSTER=$cl timeout 600 python -m pytest tests/test_w8a8.py -x -q -k "prefill or external or k3d" 2>&1 | tail -3
Q100_I8_CLUSTER=$cl timeout 120 python tools/exp_w8a8.py 2>&1 | tail -5
done
