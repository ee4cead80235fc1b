"""Cycle breakdown of the TMEM-A prefill kernel (TA_TIMING build)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "llama3-quantization_b200"))
import numpy as np, torch, q100
from q100 import _lib
dev = torch.device("cuda:0"); dt = torch.bfloat16
N, K, M = 14336, 4096, 2048
pw = q100.pack_weight((torch.randn(N, K, device=dev) * 0.02).to(dt), 4, 128)[0]
x = torch.randn(M, K, device=dev).to(dt)
fn = _lib.lib.q100_debug_tca_timing; fn.restype = C.c_int; fn.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_longlong * (8 * 148))()
for i in range(3): q100.gemm_wna16(x, pw, None, _lib.GEMM_FORCE_TC)
torch.cuda.synchronize(); fn(buf, 148)          # reset
q100.gemm_wna16(x, pw, None, _lib.GEMM_FORCE_TC); torch.cuda.synchronize()
assert fn(buf, 148) == 0
t = np.array(buf[:], dtype=np.int64).reshape(148, 8)
names = ["mma_warp_total", "wait d_empty", "wait a_full", "wait full_x", "issue+commit", "k-blocks", "dq warp0: wait full_p", "dq warp0: wait a_empty"]
for j, nm in enumerate(names):
    c = t[:, j]; print(f"{nm:22s} mean {c.mean():10.0f}  p50 {np.median(c):10.0f}  max {c.max():10.0f}")
kb = t[:, 5].mean()
print(f"per k-block (MMA warp): total {t[:,0].mean()/kb:.0f}  a_full {t[:,2].mean()/kb:.0f}  full_x {t[:,3].mean()/kb:.0f}  issue {t[:,4].mean()/kb:.0f}  d_empty/tile {t[:,1].mean()/(kb/64):.0f}")
print(f"dequant warp 0 per own k-block: wait full_p {t[:,6].mean()/(kb/2):.0f}  wait a_empty {t[:,7].mean()/(kb/2):.0f}")
