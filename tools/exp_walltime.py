"""Wall-clock (globaltimer) timeline of back-to-back decode GEMM launches in a PDL graph (SK_TIMING build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import numpy as np, torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
fn = _lib.lib.q100_debug_skinny_timing; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
def run(N, K, M, tag, pdl=True):
    pws = []
    for i in range(4):
        w = (torch.randn(N, K, device=dev) * 0.02).to(dt); pws.append(q100.pack_weight(w, 4, 128)[0]); del w
    x = torch.randn(M, K, device=dev).to(dt)
    flags = _lib.GEMM_FORCE_SKINNY | (_lib.GEMM_PDL if pdl else 0)
    for i in range(4): q100.gemm_wna16(x, pws[i % 4], None, flags)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(16): q100.gemm_wna16(x, pws[i % 4], None, flags)
    g.replay(); torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    buf = (C.c_longlong * (16 * 512 * 4))()
    assert fn(buf, -1) == 0
    t = np.array(buf[:], dtype=np.int64).reshape(16, 512, 4)
    # order launches by their earliest entry stamp
    valid = t[:, :, 0] > 0
    first = np.array([t[l, valid[l], 0].min() for l in range(16)])
    order = np.argsort(first)
    print(f"--- {tag} N={N} K={K} M={M} pdl={pdl}: per launch (us): entry spread | first entry -> first wait-done | wait-done spread | "
          f"consumer span (median / max) | resolve (max) | launch period")
    prev_end = None
    for li in order[6:10]:
        v = valid[li]
        e, w, r, d = t[li, v, 0], t[li, v, 1], t[li, v, 2], t[li, v, 3]
        period = "" if prev_end is None else f"{(e.min() - prev_start) / 1e3:6.2f}"
        print(f"  ctas {v.sum():3d} entry spread {(e.max()-e.min())/1e3:5.2f} | entry->waitdone {(w.min()-e.min())/1e3:5.2f} | waitdone spread {(w.max()-w.min())/1e3:5.2f} | "
              f"consumer med {np.median(d-w)/1e3:5.2f} max {(d-w).max()/1e3:5.2f} | resolve max {(d-r).max()/1e3:5.2f} | kernel span {(d.max()-e.min())/1e3:6.2f} | "
              f"prev_end->this first waitdone {'' if prev_end is None else f'{(w.min()-prev_end)/1e3:5.2f}'} | period {period}")
        prev_end = d.max(); prev_start = e.min()
import sys as _s
for pdl in (True, False):
    run(4096, 4096, 1, "o", pdl)
    run(14336, 4096, 1, "up", pdl)
    run(28672, 4096, 1, "gateup", pdl)
    run(4096, 14336, 1, "down", pdl)
