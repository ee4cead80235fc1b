timeout 300 python tools/trace_decode.py --layers 8 --batch 1 --out gpurun_out/tr_b1.json 2>&1 | tail -22
timeout 300 python tools/trace_decode.py --layers 8 --batch 16 --out gpurun_out/tr_b16.json 2>&1 | tail -16
rm -f gpurun_out/tr_b1.json gpurun_out/tr_b16.json
