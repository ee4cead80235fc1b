"""Kernel timeline of one decode step of the 8B model (torch.profiler / CUPTI): per-kernel durations inside a CUDA-graph
replay and the idle time between kernels.   python tools/trace_decode.py [--layers 8] [--batch 1]"""
import argparse, json, os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import torch
from torch.profiler import profile, ProfilerActivity
from q100 import llama as QL

ap = argparse.ArgumentParser()
ap.add_argument("--layers", type=int, default=8)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--out", default="gpurun_out/decode_trace.json")
a = ap.parse_args()
dev = torch.device("cuda", 0); dt = torch.float16
cfg = QL.LLAMA3_8B
model = QL.QuantLlamaModel.random_init(cfg, QL.quant_args(4, 128, 16), dev, dt, seed=0, n_layers=a.layers, fast_attention=True)
dec = QL.GraphDecoder(model, a.batch, 512)
dec.prime(256)
tok = torch.randint(0, cfg.vocab_size, (a.batch, 1)).to(dev)
dec.step(tok)
for _ in range(5): dec.step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(4): dec.step()
    torch.cuda.synchronize()
prof.export_chrome_trace(a.out)
ev = [e for e in json.load(open(a.out))["traceEvents"] if e.get("cat") == "kernel"]
ev.sort(key=lambda e: e["ts"])
# split into steps by the largest gaps
n = len(ev) // 4
step = ev[2 * n:3 * n]
t0 = step[0]["ts"]; t1 = step[-1]["ts"] + step[-1]["dur"]
busy = sum(e["dur"] for e in step)
print(f"kernels per step {n}, step span {t1 - t0:.1f} us, sum of kernel durations {busy:.1f} us")
agg = collections.OrderedDict()
for e in step:
    k = e["name"][:70]
    d = agg.setdefault(k, [0, 0.0]); d[0] += 1; d[1] += e["dur"]
for k, (c, d) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{d:9.1f} us  x{c:4d}  {d / c:7.2f} us  {k}")
# timeline of the second block
print("--- one block ---")
names = [e["name"] for e in step]
idx = [i for i, nm in enumerate(names) if "k5_rope_kv" in nm]
if len(idx) > 4:
    lo, hi = idx[2] - 1, idx[3]
    for e in step[lo:hi + 1]:
        print(f"{e['ts'] - step[lo]['ts']:8.2f} +{e['dur']:6.2f}  {e['name'][:80]}")
