#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python oracle/make_golden.py            # REF_PATH defaults to /root/reference

Imports ``quant.quantizer`` / ``quant.int_linear`` / ``quant.int_matmul`` from the
reference checkout, runs them on seeded inputs plus the SURVEY §4 edge cases,
checks ``oracle/q100_oracle.py`` against every output bit-for-bit (the script
fails if the oracle disagrees anywhere) and stores inputs and reference outputs
as raw storage bits.  The fixtures travel with the repo; the reference does not.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_PATH = os.environ.get("REF_PATH", "/root/reference")
sys.path.insert(0, REF_PATH)
sys.path.insert(0, HERE)

from quant.quantizer import UniformAffineQuantizer  # noqa: E402  (reference)
from quant.int_linear import QuantLinear            # noqa: E402  (reference)
from quant.int_matmul import QuantMatMul            # noqa: E402  (reference)
import quant.quantizer as _refq                     # noqa: E402

assert os.path.realpath(_refq.__file__).startswith(os.path.realpath(REF_PATH)), _refq.__file__

import q100_oracle as O  # noqa: E402

TDT = {"f16": torch.float16, "bf16": torch.bfloat16, "f32": torch.float32}


def t2bits(t: torch.Tensor) -> np.ndarray:
    t = t.detach().contiguous()
    if t.dtype == torch.float32:
        return t.view(torch.int32).numpy().view(np.uint32).copy()
    return t.view(torch.int16).numpy().view(np.uint16).copy()


def bits2f32(b: np.ndarray, dtype: str) -> np.ndarray:
    return O.from_bits(b, dtype)


def same_bits(a_f32: np.ndarray, ref_bits: np.ndarray, dtype: str) -> bool:
    got = O.to_bits(a_f32, dtype).reshape(ref_bits.shape)
    if np.array_equal(got, ref_bits):
        return True
    # NaNs may differ in payload/sign only
    a = O.from_bits(got, dtype)
    b = O.from_bits(ref_bits, dtype)
    return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))) and
                np.array_equal(np.signbit(a[~np.isnan(a)]), np.signbit(b[~np.isnan(b)])))


def ref_codes(q: UniformAffineQuantizer, x: torch.Tensor, g) -> torch.Tensor:
    """Restate quantizer.py:103-106 with the reference's own scale/zp."""
    xx = x
    if q.deficiency > 0:
        xx = torch.cat((xx, torch.zeros((xx.shape[0], q.deficiency), dtype=xx.dtype)), dim=1)
    if g:
        xx = xx.reshape(-1, g)
    v = xx / q.scale
    xi = (v.round() - v) + v            # round_ste forward value, quantizer.py:19
    if q.round_zero_point is not None:
        xi = xi.add(q.round_zero_point)
    return xi.clamp(q.qmin, q.qmax)


def run_weight_case(name, w_t, dtype, n_bits, group, symmetric=False, disable_zp=False,
                    lwc=False, lwc_seed=None):
    q = UniformAffineQuantizer(n_bits=n_bits, symmetric=symmetric, per_channel_axes=[0],
                               dynamic_method="per_channel", group_size=group, shape=w_t.shape,
                               lwc=lwc, disable_zero_point=disable_zp)
    up = low = None
    if lwc:
        gen = torch.Generator().manual_seed(lwc_seed)
        with torch.no_grad():
            q.upbound_factor.copy_(4.0 + torch.randn(q.upbound_factor.shape, generator=gen))
            q.lowbound_factor.copy_(4.0 + torch.randn(q.lowbound_factor.shape, generator=gen))
        q = q.to(w_t.dtype)
        up = torch.sigmoid(q.upbound_factor.detach())
        low = torch.sigmoid(q.lowbound_factor.detach())
    with torch.no_grad():
        deq = q(w_t)
        codes = ref_codes(q, w_t, group)
    rec = {
        "dtype": dtype, "n_bits": n_bits, "group": group or 0, "symmetric": int(symmetric),
        "disable_zp": int(disable_zp), "lwc": int(lwc), "deficiency": int(q.deficiency),
        "w": t2bits(w_t), "dequant": t2bits(deq), "scales": t2bits(q.scale),
        "codes": np.nan_to_num(codes.float().numpy(), nan=-32768).astype(np.int16),   # NaN code -> -32768
    }
    if q.round_zero_point is not None:
        rec["zeros"] = t2bits(q.round_zero_point)
    if lwc:
        rec["lwc_up"] = t2bits(up)
        rec["lwc_low"] = t2bits(low)

    # ---- pin the oracle against the reference, bit for bit
    w32 = bits2f32(rec["w"], dtype).reshape(w_t.shape)
    if q.deficiency > 0:   # quantizer.py:95-97,136-139 zero-pad
        w32 = np.concatenate([w32, np.zeros((w32.shape[0], q.deficiency), np.float32)], axis=1)
    o = O.quantize_groupwise(
        w32, dtype, group or 0, n_bits, symmetric, disable_zp,
        bits2f32(rec["lwc_up"], dtype) if lwc else None,
        bits2f32(rec["lwc_low"], dtype) if lwc else None)
    od = o["dequant"][:, : w_t.shape[1]] if q.deficiency > 0 else o["dequant"]
    ok = same_bits(od, rec["dequant"], dtype)
    ok &= same_bits(o["scales"], rec["scales"], dtype)
    if "zeros" in rec:
        ok &= same_bits(o["zeros"], rec["zeros"], dtype)
    oc = np.nan_to_num(o["codes"], nan=-32768).astype(np.int16).reshape(rec["codes"].shape)
    rc = rec["codes"]
    ok &= bool(np.array_equal(oc, np.nan_to_num(codes.float().numpy(), nan=-32768).astype(np.int16).reshape(rc.shape)))
    if not ok:
        raise SystemExit(f"ORACLE MISMATCH vs reference in case {name}")
    return rec


def rand_w(rows, cols, dtype, seed, std=0.02):
    gen = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, cols, generator=gen) * std).to(TDT[dtype])


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.manual_seed(0)
    cases = {}

    # ---- 1. group-wise weights: W{2,3,4,8} x {f16,bf16,f32} x {asym, sym, sym+no-zp, asym+no-zp}
    seed = 100
    for dtype in O.DTYPES:
        for n_bits in (2, 3, 4, 8):
            for mode, (sym, nozp) in {"asym": (False, False), "sym": (True, False),
                                      "symnozp": (True, True), "asymnozp": (False, True)}.items():
                seed += 1
                w = rand_w(12, 512, dtype, seed)
                cases[f"w_{dtype}_w{n_bits}_g128_{mode}"] = run_weight_case(
                    f"w_{dtype}_w{n_bits}_{mode}", w, dtype, n_bits, 128, sym, nozp)

    # ---- 2. other group sizes, per-channel (group_size=None)
    for dtype in O.DTYPES:
        for g in (32, 64, 256, None):
            seed += 1
            w = rand_w(16, 512, dtype, seed)
            cases[f"w_{dtype}_w4_g{g or 0}_asym"] = run_weight_case(f"g{g}", w, dtype, 4, g)
        seed += 1
        w = rand_w(8, 4096, dtype, seed)
        cases[f"w_{dtype}_w8_g0_asym_k4096"] = run_weight_case("perchan", w, dtype, 8, None)

    # ---- 3. LWC (post-sigmoid multipliers stored; quantizer.py:143-145)
    for dtype in O.DTYPES:
        for n_bits, sym in ((4, False), (3, False), (2, True), (4, True)):
            seed += 1
            w = rand_w(16, 512, dtype, seed)
            cases[f"w_{dtype}_w{n_bits}_g128_lwc_{'sym' if sym else 'asym'}"] = run_weight_case(
                "lwc", w, dtype, n_bits, 128, sym, False, lwc=True, lwc_seed=seed)
        seed += 1
        w = rand_w(16, 512, dtype, seed)
        cases[f"w_{dtype}_w4_g0_lwc_asym"] = run_weight_case(
            "lwc_perchan", w, dtype, 4, None, False, False, lwc=True, lwc_seed=seed)
        # deficiency: K % g != 0 needs lwc + symmetric (quantizer.py:70-75)
        seed += 1
        w = rand_w(8, 200, dtype, seed)
        cases[f"w_{dtype}_w4_g128_lwc_sym_deficient"] = run_weight_case(
            "deficient", w, dtype, 4, 128, True, False, lwc=True, lwc_seed=seed)

    # ---- 4. edge cases of SURVEY §4 (one row per phenomenon, g = 128)
    for dtype in O.DTYPES:
        rows = []
        base = torch.randn(128, generator=torch.Generator().manual_seed(7)) * 0.02
        rows.append(base.abs() + 0.5)                       # all positive  -> negative zp
        rows.append(-(base.abs() + 0.5))                    # all negative  -> zp > qmax
        rows.append(torch.full((128,), 0.3))                # constant      -> CLIPMIN scale, zp clamp -1e4
        rows.append(torch.zeros(128))                       # zeros
        rows.append(torch.full((128,), -0.3))               # constant negative -> zp clamp +1e4
        r = base.clone(); r[5] = 3.0e4; rows.append(r)      # huge outlier (scale near clamp / f16 range)
        r = base.clone(); r[0] = 6.0e4; r[1] = -6.0e4; rows.append(r)   # range overflows f16 -> inf
        rows.append(base * 1e-4)                            # tiny values   -> CLIPMIN scale
        r = torch.arange(128, dtype=torch.float32) / 127.0 * 15.0 - 7.5
        rows.append(r * 0.01)                               # exact .5 ties for round-half-even
        rows.append(torch.arange(128, dtype=torch.float32) * 0.5)       # ties again, wide range
        r = base.clone(); r[3] = float("nan"); rows.append(r)           # NaN poisons the group
        r = base.clone(); r[9] = float("inf"); rows.append(r)           # +inf
        rows.append(torch.full((128,), 1.0))                # constant 1.0: x/CLIPMIN overflows f16 -> round_ste NaN
        r = base.clone(); r[:] = 0.0; r[0] = -1e-3; rows.append(r)      # -0.0 / +0.0 handling of round_ste
        w = torch.stack(rows).to(TDT[dtype])
        for n_bits in (2, 3, 4, 8):
            cases[f"edge_{dtype}_w{n_bits}_asym"] = run_weight_case("edge", w, dtype, n_bits, 128)
        cases[f"edge_{dtype}_w4_sym"] = run_weight_case("edge_sym", w, dtype, 4, 128, True, False)
        cases[f"edge_{dtype}_w4_symnozp"] = run_weight_case("edge_symnozp", w, dtype, 4, 128, True, True)

    # ---- 5. per-token activations through QuantLinear.act_quantizer / QuantMatMul (3-D / 4-D input)
    act = {}
    for dtype in O.DTYPES:
        for a_bits in (4, 8):
            for shape in ((2, 5, 512), (1, 4, 3, 128), (3, 4096), (2, 1000)):
                seed += 1
                gen = torch.Generator().manual_seed(seed)
                x = (torch.randn(*shape, generator=gen) * 1.5).to(TDT[dtype])
                mm = QuantMatMul({"n_bits": a_bits, "per_channel_axes": [], "symmetric": False,
                                  "dynamic_method": "per_token"},
                                 {"n_bits": a_bits, "per_channel_axes": [], "symmetric": False,
                                  "dynamic_method": "per_token"}, matmul_func=torch.matmul)
                mm.set_quant_state(False, True)
                with torch.no_grad():
                    y = mm.quant_x1(x)
                q = mm.x1_quantizer
                rec = {"dtype": dtype, "n_bits": a_bits, "x": t2bits(x), "dequant": t2bits(y),
                       "scales": t2bits(q.scale), "zeros": t2bits(q.round_zero_point)}
                o = O.fakequant_rows(bits2f32(rec["x"], dtype).reshape(shape), dtype, a_bits)
                ok = same_bits(o["dequant"], rec["dequant"], dtype) and \
                    same_bits(o["scales"], rec["scales"].reshape(-1, 1), dtype) and \
                    same_bits(o["zeros"], rec["zeros"].reshape(-1, 1), dtype)
                if not ok:
                    raise SystemExit(f"ORACLE MISMATCH (activation) {dtype} a{a_bits} {shape}")
                act[f"act_{dtype}_a{a_bits}_{'x'.join(map(str, shape))}"] = rec

    # ---- 6. QuantLinear.forward (weight re-quantised each call; optional act quant): tolerance targets
    lin = {}
    for dtype in ("f16", "bf16", "f32"):
        for (w_bits, a_bits, group) in ((4, 16, 128), (3, 16, 128), (2, 16, 128), (8, 8, None), (4, 4, None)):
            seed += 1
            gen = torch.Generator().manual_seed(seed)
            lin_mod = torch.nn.Linear(512, 48, bias=True)
            with torch.no_grad():
                lin_mod.weight.copy_(torch.randn(48, 512, generator=gen) * 0.02)
                lin_mod.bias.copy_(torch.randn(48, generator=gen) * 0.1)
            lin_mod = lin_mod.to(TDT[dtype])
            ql = QuantLinear(lin_mod,
                             {"n_bits": w_bits, "per_channel_axes": [0], "symmetric": False,
                              "dynamic_method": "per_channel", "group_size": group, "lwc": False,
                              "disable_zero_point": False},
                             {"n_bits": a_bits, "per_channel_axes": [], "symmetric": False,
                              "dynamic_method": "per_token"})
            ql.set_quant_state(True, a_bits < 16)
            x = (torch.randn(2, 7, 512, generator=gen)).to(TDT[dtype])
            with torch.no_grad():
                y = ql(x)
            rec = {"dtype": dtype, "w_bits": w_bits, "a_bits": a_bits, "group": group or 0,
                   "x": t2bits(x), "w": t2bits(lin_mod.weight), "bias": t2bits(lin_mod.bias),
                   "y": t2bits(y)}
            yo = O.quant_linear(bits2f32(rec["x"], dtype).reshape(2, 7, 512),
                                bits2f32(rec["w"], dtype).reshape(48, 512),
                                bits2f32(rec["bias"], dtype), dtype, group or 0, w_bits, a_bits)
            yr = bits2f32(rec["y"], dtype).reshape(yo.shape)
            err = float(np.max(np.abs(yo - yr)))
            ref_mag = float(np.max(np.abs(yr)))
            tol = {"f16": 2e-3, "bf16": 1.6e-2, "f32": 1e-5}[dtype] * max(ref_mag, 1.0)
            if err > tol:
                raise SystemExit(f"ORACLE linear mismatch {dtype} w{w_bits}a{a_bits}: {err} > {tol}")
            lin[f"lin_{dtype}_w{w_bits}a{a_bits}_g{group or 0}"] = rec

    def dump(fname, d):
        flat = {}
        for cname, rec in d.items():
            for k, v in rec.items():
                flat[f"{cname}/{k}"] = np.asarray(v)
        np.savez_compressed(os.path.join(out_dir, fname), **flat)
        print(f"{fname}: {len(d)} cases, {os.path.getsize(os.path.join(out_dir, fname)) / 1024:.0f} KiB")

    dump("weights.npz", cases)
    dump("activations.npz", act)
    dump("linear.npz", lin)
    print("oracle == reference on all cases; torch", torch.__version__)


if __name__ == "__main__":
    main()
