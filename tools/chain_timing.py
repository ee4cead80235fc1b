"""Per-CTA cycle breakdown of the LAST GEMM phase of a small chain vs the same GEMM as a stand-alone launch
(needs the SK_TIMING build: Q100_VARIANT=TIMING Q100_NVCC_FLAGS=-DSK_TIMING python tools/build_ext.py;
run with Q100_LIB=llama3-quantization_b200/q100/libq100_TIMING.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import numpy as np, torch, q100
from q100 import _lib, chain as QC
dev = torch.device("cuda:0"); dt = torch.float16
fn = _lib.lib.q100_debug_skinny_timing; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
names = ["consumer_total", "stage_x", "wait_full", "load+compute", "finish", "blocks", "prod_wait_empty", "prod_total"]
def show(tag, n=444):
    torch.cuda.synchronize()
    buf = (C.c_longlong * (8 * n))()
    assert fn(buf, n) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(n, 8)
    print(f"--- {tag}: per-CTA cycles (mean / p50 / max)")
    for j, nm in enumerate(names):
        c = t[:, j]; print(f"  {nm:16s} {c.mean():10.0f} {np.median(c):10.0f} {c.max():10.0f}")
    print(f"  per block: compute {t[:,3].sum()/max(t[:,5].sum(),1):.0f} cyc, wait {t[:,2].sum()/max(t[:,5].sum(),1):.0f} cyc")
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1
H, I = 4096, 14336
def pw(n, k): return q100.pack_weight((torch.randn(n, k, device=dev) * 0.02).to(dt), 4, 128)[0]
wo, wgu, wd = pw(H, H), pw(2 * I, H), pw(H, I)
x = torch.randn(M, H, device=dev).to(dt)
for i in range(4): q100.gemm_wna16(x, wgu, None, _lib.GEMM_FORCE_SKINNY)
show(f"stand-alone gate|up M={M}")
att = torch.randn(M, H, device=dev).to(dt); h = torch.randn(M, H, device=dev).to(dt)
nw = torch.ones(H, device=dev, dtype=dt); gu = torch.zeros(M, 2 * I, device=dev, dtype=dt)
sa = torch.zeros(QC.ssq_parts(H), M, device=dev); sb = torch.zeros_like(sa)
ch = QC.Chain(M, 128, 4, False, dt, dev)
ch.gemm(att, wo, h, epi_mode=_lib.EPI_ADD_SSQ, residual=h, ssq_out=sb)
ch.gemm(h, wgu, gu, x_mode=_lib.X_NORM, norm_weight=nw, ssq_in=sb, eps=1e-5)
ch.build()
for i in range(4): ch.run()
show(f"chain [o, gate|up]: gate|up phase M={M}")
ch2 = QC.Chain(M, 128, 4, False, dt, dev)
ch2.gemm(att, wo, h, epi_mode=_lib.EPI_ADD_SSQ, residual=h, ssq_out=sb)
ch2.gemm(h, wgu, gu, x_mode=_lib.X_NORM, norm_weight=nw, ssq_in=sb, eps=1e-5)
ch2.gemm(gu, wd, h, x_mode=_lib.X_GATE_UP, epi_mode=_lib.EPI_ADD_SSQ, residual=h, ssq_out=sa)
ch2.build()
for i in range(4): ch2.run()
show(f"chain [o, gate|up, down]: down phase M={M}")
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for c, tag in ((ch, "o+gateup"), (ch2, "o+gateup+down")):
    s.record()
    for i in range(20): c.run()
    e.record(); e.synchronize()
    print(tag, "us per launch", s.elapsed_time(e) / 20 * 1e3)
