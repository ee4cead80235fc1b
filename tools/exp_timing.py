"""Per-CTA cycle breakdown of the decode kernel (needs the SK_TIMING build: Q100_LIB=.../libq100_TIMING.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import numpy as np, torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
def run(N, K, M, tag):
    pws = []
    for i in range(3):
        w = (torch.randn(N, K, device=dev) * 0.02).to(dt); pws.append(q100.pack_weight(w, 4, 128)[0]); del w
    x = torch.randn(M, K, device=dev).to(dt)
    for i in range(6): q100.gemm_wna16(x, pws[i % 3], None, _lib.GEMM_FORCE_SKINNY)
    torch.cuda.synchronize()
    n = 444
    buf = (C.c_longlong * (8 * n))()
    fn = _lib.lib.q100_debug_skinny_timing; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
    assert fn(buf, n) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(n, 8)
    names = ["consumer_total", "stage_x", "wait_full", "load+compute", "finish", "blocks", "prod_wait_empty", "prod_total"]
    print(f"--- {tag} N={N} K={K} M={M}: per-CTA cycles (mean / p50 / max)")
    for j, nm in enumerate(names):
        c = t[:, j]; print(f"  {nm:16s} {c.mean():10.0f} {np.median(c):10.0f} {c.max():10.0f}")
    print(f"  per block: compute {t[:,3].sum()/t[:,5].sum():.0f} cyc, wait {t[:,2].sum()/t[:,5].sum():.0f} cyc")
run(57344, 8192, 1, "gateup70")
run(14336, 4096, 1, "up")
run(14336, 4096, 16, "up")
