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
    for rows, cols in ((7, 256), (128, 64), (300, 192)):
        codes = rng.randint(qmin, qmax + 1, size=(rows, cols)).astype(np.float32)
        words = O.pack_codes(codes, n_bits, nozp)
        assert words.shape == O.packed_shape(rows, cols, n_bits) and words.dtype == np.uint32
        assert np.array_equal(O.unpack_codes(words, rows, cols, n_bits, nozp), codes)


def test_pack_layout_known_answer():
    # W4: element e of a word sits at bit 4*(e>>1) + 16*(e&1); a row of a tile is 8 words
    codes = np.zeros((130, 128), np.float32)
    codes[0, :8] = np.arange(8)
    codes[1, 64:72] = 15                      # row 1, second k-tile
    codes[129, 8:16] = np.arange(8)           # second n-tile, row 1, word 1
    w = O.pack_codes(codes, 4)
    assert w.shape == (2, 2, 1024)
    assert int(w[0, 0, 0]) == 0x75316420
    assert int(w[0, 1, 8]) == 0xFFFFFFFF and int(w[0, 0, 8]) == 0
    assert int(w[1, 0, 8 + 1]) == 0x75316420
    # W2: 16 codes per word, 4 words per row
    c2 = np.zeros((1, 64), np.float32); c2[0, :2] = (1, 2); c2[0, 16] = 3
    w2 = O.pack_codes(c2, 2)
    assert w2.shape == (1, 1, 512) and int(w2[0, 0, 0]) == (1 | (2 << 16)) and int(w2[0, 0, 1]) == 3
    # W3: plane of low 2 bits (128 rows x 4 words) then plane of high bits (128 rows x 2 words)
    c3 = np.zeros((2, 64), np.float32); c3[1, :2] = (5, 6)
    w3 = O.pack_codes(c3, 3)
    assert w3.shape == (1, 1, 768)
    assert int(w3[0, 0, 4]) == (1 | (2 << 16)) and int(w3[0, 0, 512 + 2]) == (1 | (1 << 16))


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


def test_port_linear_hook_is_the_same_product():
    """``llama_port.block_forward(linear=...)``: the hook bench_model.py uses to measure how far the REFERENCE's own
    perplexity moves when its linears accumulate differently (the yardstick for abits < 16).  In fp32 the two forms
    are the same product up to accumulation order."""
    import llama_port as LP
    cfg = LP.BlockCfg(128, 256, 4, 2, w_bits=4, group=64, a_bits=8)
    w = LP.fake_quant_weights(LP.random_block_weights(cfg, 7), cfg)
    x = torch.randn(1, 9, 128, generator=torch.Generator().manual_seed(0))
    calls = []

    def counting(x_, w_, b_=None):
        calls.append(tuple(w_.shape))
        return LP.linear_fp32_once(x_, w_, b_)
    a = LP.block_forward(x, w, cfg, act_quant=True)[0]
    b = LP.block_forward(x, w, cfg, act_quant=True, linear=counting)[0]
    assert len(calls) == 7
    assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)
