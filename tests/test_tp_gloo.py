"""CPU, world_size 2, gloo: the tensor-parallel host logic (shard bounds, the row-parallel
all-reduce, the sharded per-token min/max) and the bit-exactness of sharded quantisation,
checked with the oracle standing in for the kernels (no GPU here)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import q100_oracle as O
from q100 import tp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.RandomState(0)
        N, K, M = 64, 512, 3
        w = (rng.randn(N, K) * 0.02).astype(np.float16).astype(np.float32)
        x = rng.randn(M, K).astype(np.float16).astype(np.float32)
        full = O.quantize_groupwise(w, "f16", 128, 4)

        # --- row-parallel: K split on group boundaries, all-reduce of the partial outputs
        sh = tp.shard_bounds(N, K, tp.ROW, rank, world, 128)
        wl = w[sh.n0:sh.n1, sh.k0:sh.k1]
        loc = O.quantize_groupwise(wl, "f16", 128, 4)
        assert np.array_equal(loc["dequant"], full["dequant"][:, sh.k0:sh.k1])          # bit-exact
        assert np.array_equal(loc["codes"], full["codes"][:, sh.k0:sh.k1])
        xl = tp.shard_input(torch.from_numpy(x), tp.ROW, rank, world).numpy()
        y_part = torch.from_numpy(xl.astype(np.float64) @ loc["dequant"].astype(np.float64).T)
        y = tp.reduce_output(y_part, tp.ROW)
        ref = x.astype(np.float64) @ full["dequant"].astype(np.float64).T
        assert np.allclose(y.numpy(), ref, rtol=1e-12, atol=1e-12)

        # --- column-parallel: N split, no collective; gather only for the check
        shc = tp.shard_bounds(N, K, tp.COLUMN, rank, world, 128)
        locc = O.quantize_groupwise(w[shc.n0:shc.n1], "f16", 128, 4)
        assert np.array_equal(locc["dequant"], full["dequant"][shc.n0:shc.n1])
        yc = torch.from_numpy(x.astype(np.float64) @ locc["dequant"].astype(np.float64).T)
        assert tp.reduce_output(yc.clone(), tp.COLUMN).equal(yc)                          # no-op
        yg = tp.gather_columns(yc)
        assert np.allclose(yg.numpy(), ref, rtol=1e-12, atol=1e-12)

        # --- per-token activation stats across a K-sharded activation (abits < 16)
        mn, mx = tp.rowwise_minmax_allreduce(torch.from_numpy(xl))
        assert np.array_equal(mn.numpy()[:, 0], x.min(axis=1)) and np.array_equal(mx.numpy()[:, 0], x.max(axis=1))
        # ... the one-collective form the product uses (quantizer.stat_group): [rows, 2] pairs, MAX over (-min, max);
        # quantising the local slice with the global pair reproduces the slice of the unsharded result bit for bit
        xl16 = torch.from_numpy(xl).half()
        mm = torch.stack((xl16.amin(-1), xl16.amax(-1)), dim=1)
        mmg = tp.allreduce_minmax(mm)
        xf = torch.from_numpy(x).half()
        assert torch.equal(mmg[:, 0], xf.amin(-1)) and torch.equal(mmg[:, 1], xf.amax(-1))
        for bits in (8, 4):
            whole = O.quantize_groupwise(x, "f16", 0, bits)
            part = O.quantize_groupwise(xl, "f16", 0, bits, ext_minmax=mmg.float().numpy())
            k0 = rank * (K // world)
            assert np.array_equal(part["dequant"], whole["dequant"][:, k0:k0 + K // world])
            assert np.array_equal(part["scales"], whole["scales"]) and np.array_equal(part["zeros"], whole["zeros"])
        results[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_tp_world2_gloo():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        results = mgr.dict()
        procs = [ctx.Process(target=_worker, args=(r, world, port, results)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
        assert dict(results) == {0: "ok", 1: "ok"}


def test_shard_bounds():
    for world in (1, 2, 4, 8):
        for name, (N, K) in {"q_proj": (4096, 4096), "k_proj": (1024, 4096), "up_proj": (14336, 4096),
                             "o_proj": (4096, 4096), "down_proj": (4096, 14336)}.items():
            mode = tp.LLAMA_PLAN[name]
            cover_n, cover_k = 0, 0
            for r in range(world):
                s = tp.shard_bounds(N, K, mode, r, world, 128)
                n, k = s.shape
                assert n % 8 == 0 and k % 128 == 0
                cover_n += n if mode == tp.COLUMN else 0
                cover_k += k if mode == tp.ROW else 0
            assert cover_n == (N if mode == tp.COLUMN else 0) and cover_k == (K if mode == tp.ROW else 0)
    # per-channel weights: K is cut on the packed layout's 128-k blocks; the caller supplies all-reduced row statistics
    sh = tp.shard_bounds(4096, 4096, tp.ROW, 1, 2, 0)
    assert (sh.k0, sh.k1) == (2048, 4096)
    with pytest.raises(ValueError):
        tp.shard_bounds(4096, 4096 + 128, tp.ROW, 0, 8, 128)
    with pytest.raises(ValueError):
        tp.shard_bounds(4096, 4096, tp.COLUMN, 3, 2)
