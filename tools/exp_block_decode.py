"""The four decode launches of a block (q|k|v, o, gate|up, down; W4 g128, M = 1) in one PDL graph, three weight sets."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
SH = [(6144, 4096), (4096, 4096), (28672, 4096), (4096, 14336)]
sets = [[q100.pack_weight((torch.randn(n, k, device=dev) * 0.02).to(dt), 4, 128)[0] for n, k in SH] for _ in range(3)]
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
xs = [torch.randn(M, k, device=dev).to(dt) for _, k in SH]
fl = _lib.GEMM_PDL
def step(i):
    for j in range(4): q100.gemm_wna16(xs[j], sets[i % 3][j], None, fl)
for i in range(3): step(i)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(6): step(i)
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10): g.replay()
b.record(); b.synchronize()
print(f"M={M} split_units={os.environ.get('Q100_SK_SPLIT_UNITS','-')} max_split={os.environ.get('Q100_SK_MAX_SPLIT','-')}: {a.elapsed_time(b) / 60 * 1e3:.2f} us per block")
