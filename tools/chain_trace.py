#!/usr/bin/env python
"""Q100_CHAIN_TRACE=1 python tools/chain_trace.py [--bsz 1] [--layers 32]: per-phase timeline of one decode-chain launch
(globaltimer stamps the kernel records at every grid barrier): work time of each phase (last arrival - previous last
release), arrival skew (last - first arrival) and release latency (last release - last arrival), summed per phase kind."""
import argparse, ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
os.environ["Q100_CHAIN_TRACE"] = "1"
import torch
from q100 import llama as QL, _lib

ap = argparse.ArgumentParser()
ap.add_argument("--bsz", type=int, default=1)
ap.add_argument("--layers", type=int, default=32)
ap.add_argument("--ctx", type=int, default=256)
a = ap.parse_args()
dev = torch.device("cuda:0")
m = QL.QuantLlamaModel.random_init(QL.LLAMA3_8B, QL.quant_args(4, 128, 16), dev, torch.float16, n_layers=a.layers,
                                   fast_attention=True)
d = QL.GraphDecoder(m, a.bsz, a.ctx + 64, use_graph=True, chain=True)
d.prime(a.ctx)
n = len(d.chain.phases)
fn = _lib.lib.q100_debug_chain_trace
fn.restype = C.c_int
fn.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_uint64), C.c_void_p]
buf = (C.c_uint64 * (3 * n))()
st = torch.cuda.current_stream().cuda_stream
fn(d.chain.desc.data_ptr(), n, buf, st)            # arm
for _ in range(3):
    d.step()
torch.cuda.synchronize()
fn(d.chain.desc.data_ptr(), n, buf, st)            # arm again, then ONE traced launch
d.step()
torch.cuda.synchronize()
fn(d.chain.desc.data_ptr(), n, buf, st)
t = [[buf[3 * i + j] for j in range(3)] for i in range(n - 1)]
kinds = ["qkv", "attn", "o", "gateup", "down"]
agg = {k: [0.0, 0.0, 0.0, 0] for k in kinds}
prev_rel = None
rows = []
for i, (first, last, rel) in enumerate(t):
    k = kinds[i % len(kinds)]
    work = (last - prev_rel) / 1e3 if prev_rel is not None else float("nan")
    skew, lat = (last - first) / 1e3, (rel - last) / 1e3
    rows.append((i, k, work, skew, lat))
    if prev_rel is not None:
        ag = agg[k]; ag[0] += work; ag[1] += skew; ag[2] += lat; ag[3] += 1
    prev_rel = rel
print(f"{n} phases, batch {a.bsz}; total {(t[-1][2] - t[0][0]) / 1e3:.1f} us from the first arrival to the last release")
print("first 14 phases: idx kind work_us skew_us release_us")
for r in rows[:14]:
    print("  %3d %-8s %7.2f %7.2f %7.2f" % r)
print("mean per phase kind (us): work = last arrival - previous last release; skew = last - first arrival; release latency")
for k in kinds:
    w, s, l, c = agg[k]
    if c:
        print(f"  {k:8s} work {w / c:7.2f}  skew {s / c:7.2f}  release {l / c:6.2f}   (n={c})")
tot = sum(v[0] + v[2] for v in agg.values())
print(f"sum of work + release over the block phases: {tot / max(agg['o'][3], 1):.1f} us per block")

tp = os.environ.get("Q100_CHAIN_TRACE_PHASE")
if tp is not None:
    import numpy as np
    fn2 = _lib.lib.q100_debug_chain_cta_trace
    fn2.restype = C.c_int
    fn2.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]
    b2 = (C.c_uint64 * (512 * 8))()
    fn2(d.chain.desc.data_ptr(), b2, st)
    t2 = np.array(b2[:], dtype=np.int64).reshape(512, 8)
    act = t2[:, 5] > 0                       # CTAs that ran the GEMM share of the phase
    base = t2[t2[:, 0] > 0, 0].min()
    names = ["phase start", "args loaded", "rstd done", "x staged", "blocks done", "fix-up done", "arrive"]
    red = t2[act, 7]
    fix = (t2[act, 5] - t2[act, 4]) / 1e3
    order = np.argsort(-fix)[:8]
    poll = (t2[act, 1] - t2[act, 4]) / 1e3
    print("  slowest fix-ups (us), of which polling + summing (us), polling rounds:",
          [(round(float(fix[i]), 2), round(float(poll[i]), 2), int(red[i])) for i in order])
    print(f"per-CTA stamps of phase {tp} ({kinds[int(tp) % len(kinds)]}), us after the first CTA entered; {int(act.sum())} CTAs worked")
    for j, nm in enumerate(names):
        col = t2[act, j] if j in (2, 3, 4, 5) else t2[t2[:, j] > 0, j]
        if len(col):
            col = (col - base) / 1e3
            print(f"  {nm:12s} min {col.min():7.2f}  median {np.median(col):7.2f}  max {col.max():7.2f}")
