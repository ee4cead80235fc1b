"""CPU: the numpy oracle and the torch port against the fixtures generated from the
unmodified reference (oracle/make_golden.py).  Bit-exact."""
import numpy as np
import pytest
import torch

import q100_oracle as O
import torch_port as TP
from conftest import TDT, assert_same_bits, bits_to_torch, load_golden, torch_to_bits

W = load_golden("weights")
A = load_golden("activations")
L = load_golden("linear")


def _w_inputs(rec):
    dtype = str(rec["dtype"])
    w = O.from_bits(rec["w"], dtype).reshape(rec["w"].shape)
    d = int(rec["deficiency"])
    if d:
        w = np.concatenate([w, np.zeros((w.shape[0], d), np.float32)], axis=1)
    up = O.from_bits(rec["lwc_up"], dtype) if int(rec["lwc"]) else None
    low = O.from_bits(rec["lwc_low"], dtype) if int(rec["lwc"]) else None
    return dtype, w, d, up, low


@pytest.mark.parametrize("case", sorted(W))
def test_oracle_weights(case):
    rec = W[case]
    dtype, w, d, up, low = _w_inputs(rec)
    o = O.quantize_groupwise(w, dtype, int(rec["group"]), int(rec["n_bits"]), bool(rec["symmetric"]),
                             bool(rec["disable_zp"]), up, low)
    deq = o["dequant"][:, : rec["w"].shape[1]]
    assert_same_bits(O.to_bits(deq, dtype), rec["dequant"], dtype, "dequant")
    assert_same_bits(O.to_bits(o["scales"], dtype), rec["scales"], dtype, "scales")
    if "zeros" in rec:
        assert_same_bits(O.to_bits(o["zeros"], dtype), rec["zeros"], dtype, "zeros")
    else:
        assert o["zeros"] is None
    codes = np.nan_to_num(o["codes"], nan=-32768).astype(np.int16).reshape(rec["codes"].shape)
    assert np.array_equal(codes, rec["codes"])


@pytest.mark.parametrize("case", sorted(A))
def test_oracle_activations(case):
    rec = A[case]
    dtype = str(rec["dtype"])
    x = O.from_bits(rec["x"], dtype).reshape(rec["x"].shape)
    o = O.fakequant_rows(x, dtype, int(rec["n_bits"]))
    assert_same_bits(O.to_bits(o["dequant"], dtype), rec["dequant"], dtype, "dequant")
    assert_same_bits(O.to_bits(o["scales"], dtype), rec["scales"], dtype, "scales")
    assert_same_bits(O.to_bits(o["zeros"], dtype), rec["zeros"], dtype, "zeros")


@pytest.mark.parametrize("case", sorted(c for c in W if "lwc" not in c))
def test_torch_port_weights(case):
    """The torch port (used as the timed CPU baseline) reproduces the reference bits."""
    rec = W[case]
    dtype = str(rec["dtype"])
    w = bits_to_torch(rec["w"], dtype)
    with torch.no_grad():
        deq, scale, zp = TP.fake_quant(w, int(rec["n_bits"]), int(rec["group"]) or None,
                                       bool(rec["symmetric"]), bool(rec["disable_zp"]))
    assert_same_bits(torch_to_bits(deq), rec["dequant"], dtype, "dequant")
    assert_same_bits(torch_to_bits(scale), rec["scales"], dtype, "scales")
    if zp is not None:
        assert_same_bits(torch_to_bits(zp), rec["zeros"], dtype, "zeros")


@pytest.mark.parametrize("case", sorted(L))
def test_oracle_linear(case):
    rec = L[case]
    dtype = str(rec["dtype"])
    x = O.from_bits(rec["x"], dtype).reshape(rec["x"].shape)
    w = O.from_bits(rec["w"], dtype).reshape(rec["w"].shape)
    b = O.from_bits(rec["bias"], dtype)
    y = O.quant_linear(x, w, b, dtype, int(rec["group"]), int(rec["w_bits"]), int(rec["a_bits"]))
    yr = O.from_bits(rec["y"], dtype).reshape(y.shape)
    tol = {"f16": 2e-3, "bf16": 1.6e-2, "f32": 1e-5}[dtype] * max(float(np.abs(yr).max()), 1.0)
    assert float(np.abs(y - yr).max()) <= tol


@pytest.mark.parametrize("n_bits", [2, 3, 4, 8])
@pytest.mark.parametrize("nozp", [False, True])
def test_pack_roundtrip(n_bits, nozp):
    rng = np.random.RandomState(n_bits)
    qmin, qmax = O.qrange(n_bits, nozp)
    codes = rng.randint(qmin, qmax + 1, size=(7, 256)).astype(np.float32)
    words = O.pack_codes(codes, n_bits, nozp)
    assert words.shape == (7, O.words_per_row(256, n_bits)) and words.dtype == np.uint32
    assert np.array_equal(O.unpack_codes(words, 256, n_bits, nozp), codes)


def test_pack_layout_known_answer():
    # W4: element e of a word sits at bit 4*(e>>1) + 16*(e&1)
    codes = np.tile(np.arange(8, dtype=np.float32), 4).reshape(1, 32)
    w = O.pack_codes(codes, 4)
    assert w.shape == (1, 4) and int(w[0, 0]) == 0x75316420
    # W2: 16 elements per word, pairs (e0,e1) -> bits [0,2) and [16,18)
    c2 = np.array([[1, 2] + [0] * 30], dtype=np.float32)
    assert int(O.pack_codes(c2, 2)[0, 0]) == (1 | (2 << 16))
    # W3: low 2 bits as W2 words, then the high bit as 32-per-word
    c3 = np.array([[5, 6] + [0] * 30], dtype=np.float32)
    w3 = O.pack_codes(c3, 3)
    assert w3.shape == (1, 3) and int(w3[0, 0]) == (1 | (2 << 16)) and int(w3[0, 2]) == (1 | (1 << 16))


def test_rounding_helpers():
    x = np.array([1.0 + 2 ** -11, 1.0 + 3 * 2 ** -11, 65520.0, 1e-8], np.float32)
    assert np.array_equal(O.rnd(x, "f16"), x.astype(np.float16).astype(np.float32))
    t = torch.from_numpy(x).to(torch.bfloat16).float().numpy()
    assert np.array_equal(O.rnd(x, "bf16"), t)
    for dt in O.DTYPES:
        v = O.rnd(np.linspace(-3, 3, 101).astype(np.float32), dt)
        assert np.array_equal(O.from_bits(O.to_bits(v, dt), dt), v)
    # round_ste: -0 -> +0, inf -> nan
    r = O.round_ste(np.array([-0.2, np.inf, 2.5, 3.5], np.float32), "f32")
    assert not np.signbit(r[0]) and np.isnan(r[1]) and r[2] == 2.0 and r[3] == 4.0
