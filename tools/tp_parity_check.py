"""torchrun --nproc-per-node N tools/tp_parity_check.py --config w4a16|w8a8|w4a4 [--layers 4]

Tensor-parallel parity ON THE GPUS: a random-init LLaMA3-8B-shaped block stack sharded over N ranks (q/k/v/gate/up
column shards, o/down row shards, all-reduce after the row-parallel linears; q100/tp.py, q100/llama.py) against the
single-GPU stack built from the same seeds on rank 0.

Checks (asserted, then printed as one JSON line followed by "TP PARITY OK"):
  * weights: every rank's packed codes / scales / zero-points equal the slices of the single-GPU ones (bit-exact), also
    for per-channel weights of the row-parallel layers (statistics all-reduced, quantizer.stat_group);
  * activation quantiser in front of a row-parallel layer (abits < 16): the fake-quantised K-slice, scales and
    zero-points equal the single-GPU ones bit for bit (one [rows, 2] MAX all-reduce, tp.allreduce_minmax);
  * logits of a 512-token prefill and the mean NLL (main.py:122-151) against the single-GPU run: only the fp32
    summation order at the all-reduce differs.
The reference has no tensor parallelism (parallel_utils.py places whole layers); the single-GPU result is the target.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "llama3-quantization_b200"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CONFIGS = {"w4a16": (4, 128, 16), "w3a16": (3, 128, 16), "w8a8": (8, None, 8), "w4a4": (4, None, 4)}
# |delta mean NLL| allowed between the tensor-parallel and the single-GPU run.  A16: the bar of the north star
# (0.01 ppl at WikiText2's 6.1 = 1.6e-3 in NLL).  A8 / A4: a 1-ulp difference in a GEMM output (summation order at the
# all-reduce) can move a per-token activation code by one step downstream, so the two runs are two samples of the same
# quantised model; the bound is the spread measured between accumulation orders on ONE GPU (model-level
# configurations: 3.0e-3 / 3.6e-2 against the fp32-accumulating port) with head-room.
NLL_BAR = {"w4a16": 1.6e-3, "w3a16": 1.6e-3, "w8a8": 1e-2, "w4a4": 1e-1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="w4a16", choices=sorted(CONFIGS))
    ap.add_argument("--layers", type=int, default=4)
    ap.add_argument("--seq", type=int, default=512)
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from q100 import llama as QL, tp
    wbits, group, abits = CONFIGS[a.config]
    cfg = QL.LLAMA3_8B
    args = QL.quant_args(wbits, group, abits)
    dt = torch.float16
    out = {"config": a.config, "world": world, "layers": a.layers}
    with torch.no_grad():
        m_tp = QL.QuantLlamaModel.random_init(cfg, args, dev, dt, seed=0, rank=rank, world=world, n_layers=a.layers,
                                              drop_dense=False)
        # every rank builds the single-GPU stack too (4 blocks are small): the comparison then needs no gather
        m_1 = QL.QuantLlamaModel.random_init(cfg, args, dev, dt, seed=0, rank=0, world=1, n_layers=a.layers,
                                             drop_dense=False)
        # ---- weights
        for li, (lt, l1) in enumerate(zip(m_tp.layers, m_1.layers)):
            for name in QL._LINEARS:
                mod_t = getattr(lt.self_attn, name, None) or getattr(lt.mlp, name)
                mod_1 = getattr(l1.self_attn, name, None) or getattr(l1.mlp, name)
                n, k = mod_1.weight.shape
                sh = tp.shard_bounds(n, k, tp.LLAMA_PLAN[name], rank, world, group or 0)
                want = mod_1.weight[sh.n0:sh.n1, sh.k0:sh.k1]
                assert torch.equal(mod_t.weight.view(torch.int16), want.contiguous().view(torch.int16)), \
                    f"layer {li} {name}: sharded fake-quantised weight differs from the slice of the single-GPU one"
        out["weights_bit_exact"] = True
        # ---- activation quantiser of a row-parallel layer
        if abits < 16:
            g = torch.Generator(device=dev).manual_seed(7)
            x = torch.randn(1, 64, cfg.hidden_size, generator=g, device=dev).to(dt)
            xl = tp.shard_input(x, tp.ROW, rank, world)
            qt, q1 = m_tp.layers[0].self_attn.o_proj.act_quantizer, m_1.layers[0].self_attn.o_proj.act_quantizer
            yt, y1 = qt(xl), q1(x)
            assert torch.equal(yt.view(torch.int16), tp.shard_input(y1, tp.ROW, rank, world).view(torch.int16)), \
                "sharded activation fake-quant differs from the slice of the single-GPU one"
            assert torch.equal(qt.scale.view(torch.int16), q1.scale.view(torch.int16))
            assert torch.equal(qt.round_zero_point.view(torch.int16), q1.round_zero_point.view(torch.int16))
            out["act_codes_bit_exact"] = True
        # ---- logits / NLL
        gt = torch.Generator().manual_seed(1)
        toks = torch.randint(0, cfg.vocab_size, (1, a.seq), generator=gt).to(dev)
        lg_t, lg_1 = m_tp.logits(toks).float(), m_1.logits(toks).float()
        rel = float((lg_t - lg_1).abs().max() / lg_1.abs().max())
        nll_t = float(torch.nn.functional.cross_entropy(lg_t[0, :-1], toks[0, 1:]))
        nll_1 = float(torch.nn.functional.cross_entropy(lg_1[0, :-1], toks[0, 1:]))
        out.update({"logits_rel_max_err": rel, "mean_nll_tp": nll_t, "mean_nll_single": nll_1,
                    "delta_mean_nll": nll_t - nll_1, "nll_bar": NLL_BAR[a.config]})
        assert abs(nll_t - nll_1) <= NLL_BAR[a.config], out
        if abits >= 16:
            assert rel <= 2e-2, out
        # ---- one decode step on the static cache (fused route + K7 / NCCL) against the single-GPU one
        dec_t, dec_1 = QL.GraphDecoder(m_tp, 2, 64, use_graph=False), QL.GraphDecoder(m_1, 2, 64, use_graph=False)
        tok = toks[:, :2].reshape(2, 1)
        for dec in (dec_t, dec_1):
            dec.pos.fill_(0)
        nt, n1 = dec_t.step(tok).clone(), dec_1.step(tok).clone()
        out["decode_argmax_equal"] = bool(torch.equal(nt, n1))
    ok = torch.ones(1, device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps(out), flush=True)
        print("TP PARITY OK", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
