#!/usr/bin/env python
"""Exhaustive CPU proof obligation for k1_fast's scale division (csrc/q100_quant.cu, fast_params).

For every finite non-negative fp16 / bf16 value x and every divisor d = 2^n - 1 (n = 1..8: all asymmetric and
symmetric code ranges of 2..8 bits) the kernel's
    q = x * RN(1/d);  q = fma(fma(-q, d, x), RN(1/d), q)
must round to the same 16-bit value as div.rn.f32(x, d).  fp32 fma is emulated in float64 (both products are
exact there); sums that land exactly on an fp32 rounding midpoint would be double-rounding suspects and are
re-done in rational arithmetic.  Exit status 0 iff there is no mismatch (a -0 input is excluded: a range is
never -0).
"""
import sys
from fractions import Fraction

import numpy as np
import torch


def main() -> int:
    bad_total = 0
    for dt in (torch.float16, torch.bfloat16):
        bits = np.arange(65536, dtype=np.uint16)
        x = torch.from_numpy(bits.view(np.int16)).view(dt).float().numpy()
        x = x[np.isfinite(x) & (x >= 0) & ~np.signbit(x)]
        for n in range(1, 9):
            d = float(2 ** n - 1)
            d32 = np.float32(d)
            rcp = np.float32(1.0) / d32
            ref = torch.from_numpy((x / d32).astype(np.float32)).to(dt).float().numpy()
            q = (x * rcp).astype(np.float32)
            r64 = (-q.astype(np.float64)) * d + x.astype(np.float64)
            r = r64.astype(np.float32)
            assert np.array_equal(r.astype(np.float64), r64), "remainder not exact in fp32"
            s64 = r.astype(np.float64) * np.float64(rcp) + q.astype(np.float64)
            q2 = s64.astype(np.float32)
            low = s64.view(np.uint64) & np.uint64((1 << 29) - 1)
            for i in np.nonzero(low == np.uint64(1 << 28))[0]:
                ex = Fraction(float(r[i])) * Fraction(float(rcp)) + Fraction(float(q[i]))
                cands = (np.nextafter(q2[i], np.float32(-np.inf)), q2[i], np.nextafter(q2[i], np.float32(np.inf)))
                q2[i] = min(cands, key=lambda c: abs(Fraction(float(c)) - ex))
            got = torch.from_numpy(q2).to(dt).float().numpy()
            nb = int((got.view(np.uint32) != ref.view(np.uint32)).sum())
            print(f"{str(dt):16s} d={int(d):3d}  values {x.size}  16-bit mismatches {nb}")
            bad_total += nb
    print("TOTAL mismatches", bad_total)
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
