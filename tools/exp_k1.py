"""K1 store-pattern experiment: time the quantize kernel with subsets of its small outputs disabled."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import torch, q100
from q100 import _lib
from q100._lib import lib, check
dev = torch.device("cuda:0")
N, K, g, bits = 14336, 4096, 128, 4
ws = [(torch.randn(N, K, device=dev) * 0.02).half() for _ in range(3)]
ng = N * K // g
deq = torch.empty_like(ws[0]); packed = torch.empty(q100.ops.packed_shape(N, K, bits), dtype=torch.int32, device=dev)
sc = torch.empty(ng, dtype=torch.float16, device=dev); zp = torch.empty_like(sc); meta = torch.empty(K // g, N, 2, dtype=torch.float16, device=dev)

def run(tag, d, p, s, z, m, bpe):
    def fn(i):
        check(lib.q100_quantize_groupwise(ws[i % 3].data_ptr(), 0, N, K, g, bits, 0, 0, None, None,
                                          d.data_ptr() if d is not None else None, p.data_ptr() if p is not None else None,
                                          s.data_ptr() if s is not None else None, z.data_ptr() if z is not None else None,
                                          m.data_ptr() if m is not None else None, 0, torch.cuda.current_stream().cuda_stream))
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for i in range(12): fn(i)
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): gr.replay()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 36
    print(f"{tag:40s} {ms*1e3:7.1f} us  {N*K*bpe/ms/1e6:7.0f} GB/s", flush=True)
run("fake: deq + scales + zeros", deq, None, sc, zp, None, 4.03)
run("fake: deq only", deq, None, None, None, None, 4.0)
run("pack: packed + scales + zeros + meta", None, packed, sc, zp, meta, 2.53)
run("pack: packed + meta", None, packed, None, None, meta, 2.52)
run("pack: packed + scales + zeros", None, packed, sc, zp, None, 2.52)
run("pack: packed only", None, packed, None, None, None, 2.5)
run("stats only: scales + zeros", None, None, sc, zp, None, 2.0)
