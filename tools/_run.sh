py='import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v["median_ms_per_token"],3) for k,v in d["decode"].items()})'
echo "== K4 policy (late at 3-8 rows)"; timeout 300 python tools/bench_model.py --wbits 4 --ppl-samples 0 --batches 3,4,8 2>/dev/null | python -c "$py"
echo "== Q100_K4_EARLY=1"; Q100_K4_EARLY=1 timeout 300 python tools/bench_model.py --wbits 4 --ppl-samples 0 --batches 3,4,8 2>/dev/null | python -c "$py"
