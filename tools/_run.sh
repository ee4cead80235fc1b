py='import json,sys; d=json.loads(sys.stdin.read()); print({k:round(v["median_ms_per_token"],3) for k,v in d["decode"].items()})'
echo "== default"; timeout 300 python tools/bench_model.py --wbits 4 --ppl-samples 0 --batches 1,2,4,8,16 2>/dev/null | python -c "$py"
echo "== Q100_DECODE_PDL=0"; Q100_DECODE_PDL=0 timeout 300 python tools/bench_model.py --wbits 4 --ppl-samples 0 --batches 4,8,16 2>/dev/null | python -c "$py"
echo "== Q100_BLOCK_PDL=0"; Q100_BLOCK_PDL=0 timeout 300 python tools/bench_model.py --wbits 4 --ppl-samples 0 --batches 4,8,16 2>/dev/null | python -c "$py"
